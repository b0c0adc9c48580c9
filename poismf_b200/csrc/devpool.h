// poismf_b200 — process-wide cache of device allocations.
//
// The reference's C API is stateless (src/poismf.h:226-289): every call allocates its scratch and
// frees it before returning.  On the device that costs a dozen cudaMalloc/cudaFree round trips per
// call, each 0.3 ms on a quiet host and tens of ms on a busy one (measured, profiles/r1_e2e_phases.txt)
// — more than the sweep itself.  Blocks released by a call are therefore kept here and handed to the
// next request of the same size class; nothing about the caller-visible behaviour changes.
//
//   POISMF_B200_POOL_MB   cap on cached (idle) bytes per process; 0 disables caching; default: half
//                         of the device's memory.  pmf_b200_release_cache() returns everything.
//
// Callers must have synchronised the streams that used a block before releasing it (cudaFree's
// implicit device synchronisation is NOT reproduced).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>
#include <utility>

namespace pmf {

class DevPool {
public:
    static DevPool& get() { static DevPool* p = new DevPool(); return *p; }   // never destroyed: outlives the CUDA runtime teardown

    cudaError_t alloc(void** out, size_t bytes)
    {
        *out = nullptr;
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const size_t cls = size_class(bytes);
        {
            std::lock_guard<std::mutex> g(mu_);
            auto it = idle_.find(std::make_pair(dev, cls));
            if (it != idle_.end()) {
                *out = it->second;
                idle_.erase(it);
                idle_bytes_ -= cls;
                live_[*out] = std::make_pair(dev, cls);
                return cudaSuccess;
            }
        }
        e = cudaMalloc(out, cls);
        if (e == cudaErrorMemoryAllocation) {     // give the cached blocks back and retry once
            cudaGetLastError();
            release(dev);
            e = cudaMalloc(out, cls);
        }
        if (e != cudaSuccess) { *out = nullptr; return e; }
        std::lock_guard<std::mutex> g(mu_);
        live_[*out] = std::make_pair(dev, cls);
        return cudaSuccess;
    }

    void free(void* p)
    {
        if (!p) return;
        std::pair<int, size_t> key;
        {
            std::lock_guard<std::mutex> g(mu_);
            auto it = live_.find(p);
            if (it == live_.end()) { cudaFree(p); return; }     // not ours (e.g. allocated before a cap change)
            key = it->second;
            live_.erase(it);
            if (idle_bytes_ + key.second <= cap(key.first)) {
                idle_.insert(std::make_pair(key, p));
                idle_bytes_ += key.second;
                return;
            }
        }
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != key.first) cudaSetDevice(key.first);
        cudaFree(p);
        if (cur != key.first) cudaSetDevice(cur);
    }

    // release a block for good (never cached): memory other processes have mapped through CUDA IPC
    void discard(void* p)
    {
        if (!p) return;
        {
            std::lock_guard<std::mutex> g(mu_);
            live_.erase(p);
        }
        cudaFree(p);
    }

    // cudaFree every idle block of `dev` (all devices if dev < 0); returns the bytes released
    size_t release(int dev)
    {
        std::multimap<std::pair<int, size_t>, void*> take;
        {
            std::lock_guard<std::mutex> g(mu_);
            for (auto it = idle_.begin(); it != idle_.end();) {
                if (dev < 0 || it->first.first == dev) {
                    take.insert(*it);
                    idle_bytes_ -= it->first.second;
                    it = idle_.erase(it);
                } else ++it;
            }
        }
        int cur = 0;
        cudaGetDevice(&cur);
        size_t bytes = 0;
        for (auto& kv : take) {
            cudaSetDevice(kv.first.first);
            cudaFree(kv.second);
            bytes += kv.first.second;
        }
        cudaSetDevice(cur);
        return bytes;
    }

    size_t idle_bytes()
    {
        std::lock_guard<std::mutex> g(mu_);
        return idle_bytes_;
    }

    // size classes: multiples of 1/8 of the largest power of two below the request (<= 12.5 % slack, at most 256 MB)
    static size_t size_class(size_t bytes)
    {
        if (bytes < 512) return 512;
        size_t p2 = 1;
        while ((p2 << 1) <= bytes) p2 <<= 1;
        size_t gran = p2 >> 3;
        if (gran < 512) gran = 512;
        if (gran > ((size_t)256 << 20)) gran = (size_t)256 << 20;
        return (bytes + gran - 1) / gran * gran;
    }

private:
    // cap on the idle bytes kept per device: POISMF_B200_POOL_MB, else a quarter of THAT device's memory
    // (co-resident allocators such as PyTorch's need the rest; pmf_b200_release_cache() empties the pool)
    size_t cap(int dev)     // mu_ held
    {
        auto it = cap_.find(dev);
        if (it != cap_.end()) return it->second;
        size_t c = 0;
        if (const char* e = getenv("POISMF_B200_POOL_MB")) {
            const double mb = atof(e);
            c = mb > 0 ? (size_t)(mb * 1048576.0) : 0;
        } else {
            size_t fr = 0, total = 0;
            int cur = 0;
            cudaGetDevice(&cur);
            if (cur != dev) cudaSetDevice(dev);
            if (cudaMemGetInfo(&fr, &total) != cudaSuccess) { cudaGetLastError(); total = 0; }
            if (cur != dev) cudaSetDevice(cur);
            c = total / 4;
        }
        cap_[dev] = c;
        return c;
    }

    std::mutex mu_;
    std::multimap<std::pair<int, size_t>, void*> idle_;
    struct PtrHash { size_t operator()(const void* p) const { return std::hash<const void*>()(p); } };
    std::unordered_map<void*, std::pair<int, size_t>, PtrHash> live_;
    size_t idle_bytes_ = 0;
    std::unordered_map<int, size_t> cap_;
};

template <class T> static inline cudaError_t dmalloc(T** out, size_t bytes) { return DevPool::get().alloc((void**)out, bytes); }
static inline void dfree(void* p) { DevPool::get().free(p); }

}  // namespace pmf
