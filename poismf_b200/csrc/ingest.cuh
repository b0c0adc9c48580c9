// poismf_b200 — COO triplets -> CSR + CSC on the device.
//
// The reference's front end builds both orientations on the host with SciPy
// (/root/reference/poismf/__init__.py:376-416: coo.tocsr(), coo.tocsc(), i.e. duplicates summed,
// indices sorted within each row / column) and hands them to run_poismf.  Here the triplets are
// uploaded once and both orientations are produced where the sweep needs them:
//   key = row << 32 | col  ->  stable radix sort (cub)  ->  reduce-by-key (duplicates summed)
//   -> CSR;  key' = col << 32 | row of the unique entries -> stable radix sort -> CSC;
// indptr of either = lower bound of (major << 32) in the sorted keys.
// All integer work is exact; values of duplicated entries are added in a tree order (exact for
// counts, which is what the model takes).
#pragma once
#include <cuda_runtime.h>

namespace pmf {

template <class IX>
__global__ void coo_keys_kernel(const IX* __restrict__ rows, const IX* __restrict__ cols, size_t n,
                                unsigned long long dimA, unsigned long long dimB,
                                unsigned long long* __restrict__ keys, int* __restrict__ bad)
{
    int any = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long r = (unsigned long long)rows[i], c = (unsigned long long)cols[i];
        any |= (r >= dimA) | (c >= dimB);      // negative ids of signed types wrap to huge values: caught too
        keys[i] = (r << 32) | (c & 0xffffffffULL);
    }
    if (any) atomicOr(bad, 1);
}

// (major, minor) -> (minor, major)
__global__ void swap_keys_kernel(const unsigned long long* __restrict__ in, size_t n,
                                 unsigned long long* __restrict__ out)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long v = in[i];
        out[i] = (v << 32) | (v >> 32);
    }
}

__global__ void minor_ids_kernel(const unsigned long long* __restrict__ keys, size_t n, int* __restrict__ minor)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        minor[i] = (int)(keys[i] & 0xffffffffULL);
}

// ptr[m] = number of keys whose major id is < m, for m in [0, nmajor]
__global__ void key_offsets_kernel(const unsigned long long* __restrict__ keys, size_t n, size_t nmajor,
                                   long long* __restrict__ ptr)
{
    for (size_t m = blockIdx.x * (size_t)blockDim.x + threadIdx.x; m <= nmajor; m += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long target = (unsigned long long)m << 32;
        size_t lo = 0, hi = n;
        while (lo < hi) {
            const size_t mid = lo + ((hi - lo) >> 1);
            if (keys[mid] < target) lo = mid + 1; else hi = mid;
        }
        ptr[m] = (long long)lo;
    }
}

template <class IX>
__global__ void widen_ids_kernel(const int* __restrict__ in, size_t n, IX* __restrict__ out)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (IX)in[i];
}
template <class IX>
__global__ void widen_ptr_kernel(const long long* __restrict__ in, size_t n, IX* __restrict__ out)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (IX)in[i];
}

}  // namespace pmf
