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
    } else {
        auto kern = rows_warp_kernel<real, METHOD, STRICT, CACHED>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes);
        if (e != cudaSuccess) return e;
        kern<<<persistent_grid(kern, cfg), cfg.threads, cfg.smem_bytes, cfg.stream>>>(P);
    }
    return cudaGetLastError();
}

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
