// poismf_b200 — host<->device copies of large PAGEABLE buffers.
//
// The reference's callers pass ordinary (pageable) numpy / R vectors.  cudaMemcpyAsync moves those
// through the driver's single-threaded bounce buffer at ~8-10 GB/s (measured: 490 MB of config #2
// in ~50 ms, against 9 ms from page-locked memory).  Here a few host threads copy slices of the
// buffer into their own page-locked staging blocks and enqueue the DMA from there, so that host
// copies and PCIe transfers of different slices overlap.  Page-locked (or registered) buffers are
// detected and copied directly.
//
//   POISMF_B200_COPY_THREADS   staging threads (default 6, 0 = always plain cudaMemcpyAsync)
//
// h2d(): returns once every source byte has been read (the caller may reuse `src`); the DMAs are
//        ordered in `st` like a cudaMemcpyAsync.
// d2h(): page-locked destination: asynchronous like cudaMemcpyAsync; pageable: returns once the
//        data is in `dst`.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace pmf {

class Stager {
public:
    static Stager& get() { static Stager* s = new Stager(); return *s; }

    static bool page_locked(const void* p)
    {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
    }

    cudaError_t h2d(void* dst, const void* src, size_t bytes, cudaStream_t st) { return run(dst, src, bytes, st, true); }
    cudaError_t d2h(void* dst, const void* src, size_t bytes, cudaStream_t st) { return run(dst, src, bytes, st, false); }

private:
    static constexpr size_t CHUNK = (size_t)4 << 20;      // staging block
    static constexpr size_t MIN_STAGED = (size_t)8 << 20; // below this the plain path is as fast
    static constexpr int MAX_THREADS = 16;
    struct Lane {
        char* buf[2] = {nullptr, nullptr};
        cudaEvent_t ev[2] = {nullptr, nullptr};
    };
    std::mutex mu_;               // one staged transfer at a time (the lanes are shared)
    Lane lanes_[MAX_THREADS];
    int lanes_device_ = -1;
    int nthreads_ = -1;

    int threads()
    {
        if (nthreads_ >= 0) return nthreads_;
        int n = 6;
        if (const char* e = getenv("POISMF_B200_COPY_THREADS")) n = atoi(e);
        const int hw = (int)std::thread::hardware_concurrency();
        if (hw > 0) n = std::min(n, hw);
        nthreads_ = std::max(0, std::min(n, MAX_THREADS));
        return nthreads_;
    }

    cudaError_t ensure_lanes(int dev, int T)
    {
        if (lanes_device_ >= 0 && lanes_device_ != dev) {       // events are per device: rebuild
            for (auto& l : lanes_) for (int j = 0; j < 2; j++) {
                if (l.ev[j]) { cudaEventSynchronize(l.ev[j]); cudaEventDestroy(l.ev[j]); l.ev[j] = nullptr; }
            }
        }
        lanes_device_ = dev;
        for (int t = 0; t < T; t++) for (int j = 0; j < 2; j++) {
            if (!lanes_[t].buf[j]) {
                cudaError_t e = cudaHostAlloc((void**)&lanes_[t].buf[j], CHUNK, cudaHostAllocPortable);
                if (e != cudaSuccess) return e;
            }
            if (!lanes_[t].ev[j]) {
                cudaError_t e = cudaEventCreateWithFlags(&lanes_[t].ev[j], cudaEventDisableTiming);
                if (e != cudaSuccess) return e;
            }
        }
        return cudaSuccess;
    }

    cudaError_t run(void* dst, const void* src, size_t bytes, cudaStream_t st, bool up)
    {
        const cudaMemcpyKind kind = up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        const void* host = up ? src : dst;
        const int T = threads();
        if (bytes < MIN_STAGED || T == 0 || page_locked(host)) return cudaMemcpyAsync(dst, src, bytes, kind, st);
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        std::lock_guard<std::mutex> g(mu_);
        e = ensure_lanes(dev, T);
        if (e != cudaSuccess) return e;
        const size_t nchunks = (bytes + CHUNK - 1) / CHUNK;
        std::vector<cudaError_t> errs(T, cudaSuccess);
        auto work = [&](int t) {
            cudaError_t err = cudaSetDevice(dev);
            Lane& L = lanes_[t];
            int j = 0;
            for (size_t c = t; c < nchunks && err == cudaSuccess; c += T, j ^= 1) {
                const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
                err = cudaEventSynchronize(L.ev[j]);          // the block's previous DMA has drained
                if (err != cudaSuccess) break;
                if (up) {
                    memcpy(L.buf[j], (const char*)src + off, len);
                    err = cudaMemcpyAsync((char*)dst + off, L.buf[j], len, kind, st);
                    if (err == cudaSuccess) err = cudaEventRecord(L.ev[j], st);
                } else {
                    err = cudaMemcpyAsync(L.buf[j], (const char*)src + off, len, kind, st);
                    if (err == cudaSuccess) err = cudaEventRecord(L.ev[j], st);
                    if (err == cudaSuccess) err = cudaEventSynchronize(L.ev[j]);
                    if (err == cudaSuccess) memcpy((char*)dst + off, L.buf[j], len);
                }
            }
            errs[t] = err;
        };
        std::vector<std::thread> pool;
        pool.reserve(T - 1);
        for (int t = 1; t < T; t++) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
        for (int t = 0; t < T; t++) if (errs[t] != cudaSuccess) return errs[t];
        return cudaSuccess;
    }
};

}  // namespace pmf
