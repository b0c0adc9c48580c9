// poismf_b200 — host-side launch descriptors shared by api.cu and the kernel TUs.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace pmf {

template <class real> struct SideParams;

struct LaunchCfg {
    bool block_team;     // false: one warp per row; true: one CTA per row
    bool cached;         // cg: re-use <x,F_t>, <d,F_t> in the line search (fast mode, limit_step)
    int threads;
    size_t smem_bytes;
    cudaStream_t stream;
    int needed;          // CTAs needed to give every row its own team
    int max_grid;        // upper bound on the grid (global-scratch bin)
    int num_sms;
};

template <class real> cudaError_t launch_rows_pgcg_fast(const LaunchCfg&, const SideParams<real>&);
template <class real> cudaError_t launch_rows_pgcg_strict(const LaunchCfg&, const SideParams<real>&);
template <class real> cudaError_t launch_rows_tn_fast(const LaunchCfg&, const SideParams<real>&);
template <class real> cudaError_t launch_rows_tn_strict(const LaunchCfg&, const SideParams<real>&);

}  // namespace pmf
