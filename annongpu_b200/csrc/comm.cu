// NCCL (dlopen) / callback transport behind allreduce_sum -- see comm.hpp.
#include "comm.hpp"
#include <dlfcn.h>
#include <cstdlib>

namespace angpu {

namespace {

// the subset of the NCCL ABI used here (nccl.h; stable since 2.x): opaque comm, 128-byte unique id, ncclFloat64 = 8, ncclSum = 0
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[COMM_ID_BYTES]; };
typedef int ncclResult_t;
struct Nccl {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
    const char*  (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
Nccl g_nccl;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_world = 1;
allreduce_fn g_cb = nullptr;
void* g_cb_user = nullptr;
bool g_reduce = false;
cudaStream_t g_stream = nullptr;                 // the collectives' stream (allreduce_sum)
cudaEvent_t g_ev_in = nullptr, g_ev_out = nullptr;

void nccl_load() {
    if(g_nccl.so) return;
    void* so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);      // the copy already in the process (torch)
    if(!so) so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if(!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if(!so) throw Error(std::string("angpu_comm: cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* name) { void* p = dlsym(so, name); if(!p) throw Error(std::string("angpu_comm: libnccl lacks ") + name); return p; };
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
    g_nccl.CommGetAsyncError = reinterpret_cast<decltype(g_nccl.CommGetAsyncError)>(sym("ncclCommGetAsyncError"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
    g_nccl.GetVersion = reinterpret_cast<decltype(g_nccl.GetVersion)>(sym("ncclGetVersion"));
    g_nccl.so = so;
}
void nccl_check(ncclResult_t r, const char* what) {
    if(r != 0) throw Error(std::string("NCCL ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error") + " (" + std::to_string(r) + ")");
}

} // namespace

void set_allreduce(allreduce_fn fn, void* user) { g_cb = fn; g_cb_user = user; }

void comm_unique_id(unsigned char out[COMM_ID_BYTES]) {
    nccl_load();
    ncclUniqueId id;
    nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(out, id.internal, COMM_ID_BYTES);
}
void comm_init(const unsigned char id_bytes[COMM_ID_BYTES], int rank, int world) {
    ANGPU_REQUIRE(world >= 1 && rank >= 0 && rank < world, "angpu_comm_init: 0 <= rank < world");
    ANGPU_REQUIRE(ctx().device >= 0, "angpu_comm_init: call angpu_init first");
    comm_destroy();
    g_rank = rank; g_world = world;
    if(world == 1) return;
    nccl_load();
    ncclUniqueId id;
    std::memcpy(id.internal, id_bytes, COMM_ID_BYTES);
    ANGPU_CUDA(cudaSetDevice(ctx().device));
    nccl_check(g_nccl.CommInitRank(&g_comm, world, id, rank), "ncclCommInitRank");
}
void comm_destroy() {
    if(g_comm) { cudaStreamSynchronize(stream()); if(g_stream) cudaStreamSynchronize(g_stream); g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
    g_rank = 0; g_world = 1;
}
int comm_rank() { return g_rank; }
int comm_world() { return g_world; }
bool comm_active() { return g_comm != nullptr || g_cb != nullptr; }
void set_reduce(bool on) { g_reduce = on; }
bool reduce_on() { return g_reduce && comm_active(); }

void allreduce_sum(double* dev_ptr, size_t count) {
    if(!reduce_on() || count == 0) return;
    if(g_comm) {
        // The collective runs on its own high-priority stream, fenced by events against the library stream (the arrangement of
        // torch's process group): measured at 8 GPUs, the same RING/LL all-reduce of the 524 KB packed sums took 0.12 ms longer
        // when it was enqueued on the compute stream itself.  ANGPU_COMM_STREAM=0 keeps it on the library stream.
        static const bool side = [] { const char* e = getenv("ANGPU_COMM_STREAM"); return !(e && atoi(e) == 0); }();
        cudaStream_t cs = stream();
        if(side) {
            if(!g_stream) {
                int lo = 0, hi = 0;
                ANGPU_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                ANGPU_CUDA(cudaStreamCreateWithPriority(&g_stream, cudaStreamNonBlocking, hi));
                ANGPU_CUDA(cudaEventCreateWithFlags(&g_ev_in, cudaEventDisableTiming));
                ANGPU_CUDA(cudaEventCreateWithFlags(&g_ev_out, cudaEventDisableTiming));
            }
            ANGPU_CUDA(cudaEventRecord(g_ev_in, stream()));
            ANGPU_CUDA(cudaStreamWaitEvent(g_stream, g_ev_in, 0));
            cs = g_stream;
        }
        nccl_check(g_nccl.AllReduce(dev_ptr, dev_ptr, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, g_comm, cs), "ncclAllReduce");
        if(side) {
            ANGPU_CUDA(cudaEventRecord(g_ev_out, g_stream));
            ANGPU_CUDA(cudaStreamWaitEvent(stream(), g_ev_out, 0));
        }
        ncclResult_t async = 0;                                   // failure detection: a dead peer / aborted communicator surfaces here
        nccl_check(g_nccl.CommGetAsyncError(g_comm, &async), "ncclCommGetAsyncError");
        nccl_check(async, "asynchronous error");
        return;
    }
    if(g_cb(dev_ptr, (unsigned long long)count, g_cb_user) != 0)
        throw Error("all-reduce callback failed (the host transport reported an error; partial sums were NOT reduced)");
}

} // namespace angpu
