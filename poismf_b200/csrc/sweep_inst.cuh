// poismf_b200 — explicit instantiation of the half-sweep kernels for one
// (numerics mode, solver family) pair.  Included by four tiny .cu files so that
// the strict variants can be compiled with --fmad=false and the large tncg
// kernels build in parallel with the pg/cg ones.
//   PMF_INST_STRICT : 0/1      PMF_INST_TN : 0 (pg + cg) / 1 (tncg)
#pragma once
#include "kernels.cuh"
#include "launch.h"

namespace pmf {

// Persistent grid: as many CTAs as can be resident (SMs x occupancy), never more
// than there are rows to hand out.
template <class K> static int persistent_grid(K kern, const LaunchCfg& cfg)
{
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, cfg.threads, cfg.smem_bytes) != cudaSuccess || occ < 1)
        occ = 1;
    long long g = (long long)cfg.num_sms * occ;
    if (g > cfg.needed) g = cfg.needed;
    if (g > cfg.max_grid) g = cfg.max_grid;
    return g < 1 ? 1 : (int)g;
}

template <class real, int METHOD, bool STRICT, bool CACHED>
static cudaError_t launch_one(const LaunchCfg& cfg, const SideParams<real>& P)
{
    cudaError_t e;
    if (cfg.block_team) {
        auto kern = rows_block_kernel<real, METHOD, STRICT, CACHED>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes);
        if (e != cudaSuccess) return e;
        kern<<<persistent_grid(kern, cfg), cfg.threads, cfg.smem_bytes, cfg.stream>>>(P);
    } else if (!STRICT && cfg.team_width == 8) {
        auto kern = rows_warp_kernel<real, METHOD, STRICT, CACHED, STRICT ? 32 : 8>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes);
        if (e != cudaSuccess) return e;
        kern<<<persistent_grid(kern, cfg), cfg.threads, cfg.smem_bytes, cfg.stream>>>(P);
    } else if (!STRICT && cfg.team_width == 16) {
        auto kern = rows_warp_kernel<real, METHOD, STRICT, CACHED, STRICT ? 32 : 16>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes);
        if (e != cudaSuccess) return e;
        kern<<<persistent_grid(kern, cfg), cfg.threads, cfg.smem_bytes, cfg.stream>>>(P);
    } else {
        auto kern = rows_warp_kernel<real, METHOD, STRICT, CACHED, 32>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes);
        if (e != cudaSuccess) return e;
        kern<<<persistent_grid(kern, cfg), cfg.threads, cfg.smem_bytes, cfg.stream>>>(P);
    }
    return cudaGetLastError();
}

#if !PMF_INST_STRICT
// ---- cluster-per-row launch (cudaLaunchKernelEx with a cluster dimension) ----
template <class real, int METHOD, bool CACHED>
static cudaError_t launch_gang_one(const LaunchCfg& cfg, const SideParams<real>& P)
{
    auto kern = rows_cluster_kernel<real, METHOD, CACHED>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes);
    if (e != cudaSuccess) return e;
    if (cfg.cluster > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t lc = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cfg.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    lc.blockDim = dim3(cfg.threads); lc.dynamicSmemBytes = cfg.smem_bytes; lc.stream = cfg.stream;
    lc.gridDim = dim3(cfg.cluster);   // placeholder for the occupancy query
    int nclusters = 0;
    e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &lc);
    if (e != cudaSuccess || nclusters < 1) { cudaGetLastError(); nclusters = cfg.num_sms / cfg.cluster / 2; if (nclusters < 1) nclusters = 1; }
    if (nclusters > cfg.needed) nclusters = cfg.needed;
    if (nclusters > cfg.max_grid) nclusters = cfg.max_grid;
    lc.gridDim = dim3((unsigned)(nclusters * cfg.cluster));
    return cudaLaunchKernelEx(&lc, kern, P);
}
#if PMF_INST_TN
template <class real> cudaError_t launch_gang_tn_fast(const LaunchCfg& cfg, const SideParams<real>& P)
{
    return launch_gang_one<real, M_TNCG, false>(cfg, P);
}
template cudaError_t launch_gang_tn_fast<float>(const LaunchCfg&, const SideParams<float>&);
template cudaError_t launch_gang_tn_fast<double>(const LaunchCfg&, const SideParams<double>&);
#else
template <class real> cudaError_t launch_gang_pgcg_fast(const LaunchCfg& cfg, const SideParams<real>& P)
{
    if (P.hc.method == M_PG) return launch_gang_one<real, M_PG, false>(cfg, P);
    if (cfg.cached) return launch_gang_one<real, M_CG, true>(cfg, P);
    return launch_gang_one<real, M_CG, false>(cfg, P);
}
template cudaError_t launch_gang_pgcg_fast<float>(const LaunchCfg&, const SideParams<float>&);
template cudaError_t launch_gang_pgcg_fast<double>(const LaunchCfg&, const SideParams<double>&);
#endif
#endif

#if PMF_INST_TN
#define PMF_LAUNCH_NAME_(s) launch_rows_tn_##s
#else
#define PMF_LAUNCH_NAME_(s) launch_rows_pgcg_##s
#endif
#if PMF_INST_STRICT
#define PMF_LAUNCH_NAME PMF_LAUNCH_NAME_(strict)
#else
#define PMF_LAUNCH_NAME PMF_LAUNCH_NAME_(fast)
#endif

template <class real>
cudaError_t PMF_LAUNCH_NAME(const LaunchCfg& cfg, const SideParams<real>& P)
{
    constexpr bool S = PMF_INST_STRICT != 0;
#if PMF_INST_TN
    return launch_one<real, M_TNCG, S, false>(cfg, P);
#else
    if (P.hc.method == M_PG) return launch_one<real, M_PG, S, false>(cfg, P);
    if (!S && cfg.cached) return launch_one<real, M_CG, S, true>(cfg, P);
    return launch_one<real, M_CG, S, false>(cfg, P);
#endif
}
template cudaError_t PMF_LAUNCH_NAME<float>(const LaunchCfg&, const SideParams<float>&);
template cudaError_t PMF_LAUNCH_NAME<double>(const LaunchCfg&, const SideParams<double>&);

}  // namespace pmf
