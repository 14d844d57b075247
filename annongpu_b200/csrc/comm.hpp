// Multi-GPU communicator of libangpu: one process per GPU, sums over ranks on the library's stream.
//
// The path shards with no data-path exchange (chains / basis ranges are independent); the only collectives are sums of
// partial results.  They run through NCCL, opened at run time with dlopen("libnccl.so.2") -- inside a
// Python process that already imported torch this resolves to the NCCL torch loaded, a plain C / C++ user gets the
// system library -- so libangpu has no link-time dependency on NCCL and single-GPU users never load it.
// A host callback (angpu_set_allreduce) remains as the fallback transport for hosts that own their communicator.
// The reference has no counterpart (single device, no collectives).
#pragma once
#include "runtime.hpp"

namespace angpu {

// callback transport: sum `count` doubles in place across ranks, stream-ordered on angpu's stream; non-zero = failure
typedef int (*allreduce_fn)(void* dev_ptr, unsigned long long count, void* user);
void set_allreduce(allreduce_fn fn, void* user);

constexpr int COMM_ID_BYTES = 128;                 // sizeof(ncclUniqueId)
void comm_unique_id(unsigned char out[COMM_ID_BYTES]);                    // rank 0: ncclGetUniqueId
void comm_init(const unsigned char id[COMM_ID_BYTES], int rank, int world);   // every rank: ncclCommInitRank on the current device
void comm_destroy();
int  comm_rank();
int  comm_world();                                  // 1 when no communicator / callback is installed
bool comm_active();                                 // a transport is installed (NCCL or callback)

// Collectives are issued only while the CURRENT evaluation is sharded: Ensemble::generate switches this on iff the
// ensemble's world > 1, so an ensemble that runs all chains on one rank is never summed over ranks.
void set_reduce(bool on);
bool reduce_on();                                   // comm_active() && the current evaluation is sharded
void allreduce_sum(double* dev_ptr, size_t count);  // no-op unless reduce_on(); throws angpu::Error on failure

} // namespace angpu
