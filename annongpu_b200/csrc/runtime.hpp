// Host-side runtime of libangpu: device context (one GPU, one stream per process), owning device
// buffers, and the launch geometry helpers.  Replaces the reference's Array<T> dual host/device
// buffer with explicit update_host/update_device (include/Array.hpp:35-138): results stay resident
// in HBM and cross to the host only through the C-ABI getters.
#pragma once
#include "common.cuh"
#include <vector>
#include <cstring>

namespace angpu {

struct Ctx {
    int          device = -1;
    cudaStream_t stream = nullptr;
    bool         own_stream = false;
    int          num_sms = 148;
    size_t       smem_optin = 227 * 1024;
    unsigned long long launches = 0;   // kernels launched by this library (bench.py's gpu_launches)
};
Ctx& ctx();
void ctx_init(int device);            // cudaSetDevice + stream; idempotent per device
inline cudaStream_t stream() { return ctx().stream; }
inline void count_launch(unsigned n = 1) { ctx().launches += n; }

template<typename T>
struct DevBuf {
    T*     p = nullptr;
    size_t n = 0, cap = 0;

    DevBuf() = default;
    explicit DevBuf(size_t n_) { resize(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if(this != &o) { release(); p = o.p; n = o.n; cap = o.cap; o.p = nullptr; o.n = o.cap = 0; } return *this; }
    ~DevBuf() { release(); }

    void release() { if(p) cudaFree(p); p = nullptr; n = cap = 0; }
    // grow-only: contents are NOT preserved on growth
    void resize(size_t n_) {
        if(n_ > cap) {
            if(p) { ANGPU_CUDA(cudaStreamSynchronize(stream())); cudaFree(p); p = nullptr; n = cap = 0; }
            T* q = nullptr;
            const cudaError_t e = cudaMalloc(&q, sizeof(T) * (n_ ? n_ : 1));
            if(e != cudaSuccess) {
                cudaGetLastError();                               // clear the (non-sticky) allocation error; the buffer stays empty
                throw Error(std::string("cudaMalloc of ") + std::to_string(sizeof(T) * n_) + " bytes: " + cudaGetErrorString(e));
            }
            p = q; cap = n_ ? n_ : 1;
        }
        n = n_;
    }
    void zero() { if(n) ANGPU_CUDA(cudaMemsetAsync(p, 0, sizeof(T) * n, stream())); }
    void upload(const T* src, size_t n_) {
        resize(n_);
        if(n_) ANGPU_CUDA(cudaMemcpyAsync(p, src, sizeof(T) * n_, cudaMemcpyHostToDevice, stream()));
        ANGPU_CUDA(cudaStreamSynchronize(stream()));   // src may be pageable / short-lived
    }
    void upload(const std::vector<T>& v) { upload(v.data(), v.size()); }
    void download(T* dst, size_t n_, size_t offset = 0) const {
        if(n_) ANGPU_CUDA(cudaMemcpyAsync(dst, p + offset, sizeof(T) * n_, cudaMemcpyDeviceToHost, stream()));
        ANGPU_CUDA(cudaStreamSynchronize(stream()));
    }
    std::vector<T> to_host() const { std::vector<T> v(n); download(v.data(), n); return v; }
    void copy_from(const DevBuf& o) {
        resize(o.n);
        if(n) ANGPU_CUDA(cudaMemcpyAsync(p, o.p, sizeof(T) * n, cudaMemcpyDeviceToDevice, stream()));
    }
};

inline unsigned ceil_div(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

} // namespace angpu
