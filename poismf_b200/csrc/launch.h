// poismf_b200 — host-side launch descriptors shared by api.cu and the kernel TUs.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace pmf {

template <class real> struct SideParams;

struct LaunchCfg {
    bool block_team;     // false: `team_width` lanes per row; true: one CTA per row
    int team_width;      // 8, 16 or 32 (sub-warp teams exist in fast numerics only)
    bool cached;         // cg: re-use <x,F_t>, <d,F_t> in the line search (fast mode, limit_step)
    int threads;
    size_t smem_bytes;
    cudaStream_t stream;
    int needed;          // CTAs needed to give every row its own team
    int max_grid;        // upper bound on the grid (global-scratch bin)
    int num_sms;
    int cluster;         // CTAs per row (thread-block cluster size); <= 1: no cluster
    int rt_nw = 0;       // > 0: register-tile kernel (regtile.cuh) with this many warps per row,
    int rt_tpl = 0;      //      tile rows per lane,
    int rt_nc = 0;       //      16-byte chunks per lane and tile row (k <= 16 rt_nc)
};

template <class real> cudaError_t launch_rows_pgcg_fast(const LaunchCfg&, const SideParams<real>&);
template <class real> cudaError_t launch_rows_pgcg_strict(const LaunchCfg&, const SideParams<real>&);
template <class real> cudaError_t launch_rows_tn_fast(const LaunchCfg&, const SideParams<real>&);
template <class real> cudaError_t launch_rows_tn_strict(const LaunchCfg&, const SideParams<real>&);
// cluster-per-row kernels exist in fast numerics only
template <class real> cudaError_t launch_gang_pgcg_fast(const LaunchCfg&, const SideParams<real>&);
template <class real> cudaError_t launch_gang_tn_fast(const LaunchCfg&, const SideParams<real>&);
// register-tile kernels: fast numerics, float32, w_mult == 1; pg, and cg with limit_step + cached line search
cudaError_t launch_regtile_pg(const LaunchCfg&, const SideParams<float>&);
cudaError_t launch_regtile_cg(const LaunchCfg&, const SideParams<float>&);

}  // namespace pmf
