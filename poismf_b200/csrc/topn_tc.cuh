// poismf_b200 — tensor-core candidate scorer for batched topN (sm_100a: tcgen05 + TMEM).
//
// topN's scoring step, scores[u, j] = <A[u], B[j]>, is the one dense contraction of the path
// (/root/reference/src/topN.c:215-224 does it as one GEMV per user).  Batched over users it is a
// GEMM; here it runs on the 5th-generation tensor cores in TF32:
//
//   * one CTA per 128 users x 128 items tile; both operand tiles are staged K-major in shared
//     memory in the canonical no-swizzle UMMA layout (8-row x 16-byte core matrices; for tile
//     row r and 16-byte K-chunk c the chunk sits at (c * 128 + r) * 16 bytes, i.e. stride-byte-
//     offset 128 B between 8-row groups, leading-byte-offset 2048 B between K-chunks),
//   * one elected thread issues k/8 `tcgen05.mma.cta_group::1.kind::tf32` instructions
//     (M = 128, N = 128, K = 8 each) accumulating into 128 TMEM columns, then
//     `tcgen05.commit` onto an mbarrier,
//   * the four warps read their 32 TMEM lanes back with `tcgen05.ld.32x32b.x32` and write the
//     scores.
//
// TF32 scores are only used to pick CANDIDATES: the caller keeps the best 2N+ per user, re-scores
// them exactly in FP32 (same left-to-right sums as the reference) and proves from the TF32 error
// bound that no other item can enter the top N; users for which the proof fails are redone by the
// exact scorer.  Rankings are therefore those of the exact path.
//
// The select is fused into the scorer's epilogue so that the U x n score matrix never exists
// (SURVEY.md 8d).  Two passes over the tiles, the GEMM being far cheaper than any sort of n scores:
//   pass 1 (MODE_GROUPMAX)  per user, the maximum of every group of 16 consecutive items
//           -> tau[u] = the M-th largest group maximum (radix select over n/16 values): at least M
//              items score >= tau[u], and at most 16 M of them unless scores tie at tau[u];
//   pass 2 (MODE_EMIT)      items with score >= tau[u] are appended to the user's candidate list,
// then the (<= 4096) candidates are ordered by (score desc, id asc) and the best M go to the exact
// re-scoring as before.  Excluded items are masked in the epilogue from a per-user bitmap.
#pragma once
#include "common.cuh"

namespace pmf {
namespace tc {

constexpr int TM = 128;   // users per tile  (UMMA M)
constexpr int TN = 128;   // items per tile  (UMMA N)
constexpr int GROUP = 16;        // items per group maximum
constexpr int CAND_CAP = 4096;   // candidate slots per user (>= GROUP * 256)
constexpr int CAND_TOP = 256;    // candidates handed to the exact re-scoring (>= M)
enum { MODE_SCORES = 0, MODE_GROUPMAX = 1, MODE_EMIT = 2 };

// order-preserving key of a non-negative score; masked entries (negative) sort below everything
PMF_DEVINL uint32_t score_key(float s) { return s < 0.f ? 0u : __float_as_uint(s) + 1u; }
PMF_DEVINL float key_score(uint32_t key) { return key == 0u ? -RealTraits<float>::huge() : __uint_as_float(key - 1u); }

PMF_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor (K-major, no swizzle), sm_100 format: start address >> 4 in bits
// [0,14), leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48).
PMF_DEVINL uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// Instruction descriptor for kind::tf32: D = F32 (bits [4,6) = 1), A = B = TF32 (bits [7,10), [10,13) = 2),
// both K-major (bits 15, 16 = 0), N >> 3 in [17,23), M >> 4 in [24,29).
constexpr uint32_t make_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

PMF_DEVINL void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
PMF_DEVINL void mbar_wait(uint64_t* bar, uint32_t phase)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(phase), "r"(0x989680u)
        : "memory");
}
PMF_DEVINL void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
PMF_DEVINL void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
PMF_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Stage `rows_valid` rows (row stride ldf floats, kpad <= padded k) of a factor matrix as one UMMA
// operand tile.  Eight consecutive threads write eight consecutive tile rows of one K-chunk: a
// conflict-free 128-byte shared-memory line per quarter-warp.
PMF_DEVINL void stage_operand(float4* s, const float* __restrict__ G, size_t row0, size_t rows_total, int ldf, int kpad,
                              int tile_rows)
{
    const int nkc = kpad / 4, rblocks = tile_rows / 8;
    for (int q = threadIdx.x >> 3; q < rblocks * nkc; q += blockDim.x >> 3) {
        const int rb = q % rblocks, kc = q / rblocks;
        const int row = rb * 8 + (threadIdx.x & 7);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const size_t gr = row0 + row;
        if (gr < rows_total && kc * 4 < ldf) v = __ldg(reinterpret_cast<const float4*>(G + gr * ldf + kc * 4));
        s[kc * tile_rows + row] = v;
    }
}

// What the epilogue of a tile does with its scores
struct TileOut {
    float* scores; int* ids;                 // MODE_SCORES: scores[u, j] (TF32), ids[u, j] = j
    const uint32_t* excl_bits; size_t excl_words;   // bitmap of excluded items, excl_words 32-bit words per user (or null)
    float* gmax; size_t ngroups;             // MODE_GROUPMAX: gmax[u, j / GROUP]
    const float* tau;                        // MODE_EMIT: per-user threshold ...
    float* cand_sc; int* cand_id; int* cand_cnt;   // ... and candidate lists (CAND_CAP slots per user)
};

// grid = (ceil(n/TN), ceil(U/TM)), 128 threads.
template <int MODE>
__global__ void __launch_bounds__(128) score_tiles_tf32_kernel(const float* __restrict__ Asel, int U,
                                                               const float* __restrict__ B, size_t n, int ldf,
                                                               int kpad, TileOut out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* sA = reinterpret_cast<float4*>(smem_raw);
    float4* sB = sA + (size_t)(kpad / 4) * TM;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t j0 = (size_t)blockIdx.x * TN;
    const int u0 = blockIdx.y * TM;

    if (warp == 0) {   // 128 TMEM columns (fp32 accumulators of the 128 x 128 tile)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)TN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) mbar_init(&mbar, 1);
    stage_operand(sA, Asel, (size_t)u0, (size_t)U, ldf, kpad, TM);
    stage_operand(sB, B, j0, n, ldf, kpad, TN);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (tid == 0) {
        constexpr uint32_t idesc = make_idesc_tf32(TM, TN);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int kb = 0; kb < kpad / 8; kb++) {   // K = 8 (two 16-byte chunks) per instruction
            const uint64_t da = make_smem_desc(a0 + (uint32_t)kb * 2u * TM * 16u, TM * 16u, 128u);
            const uint64_t db = make_smem_desc(b0 + (uint32_t)kb * 2u * TN * 16u, TN * 16u, 128u);
            umma_tf32(tmem, da, db, idesc, kb > 0 ? 1u : 0u);
        }
        umma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w owns TMEM lanes 32w .. 32w+31 (= tile rows), 32 columns at a time
    const int row = warp * 32 + lane;
    const int u = u0 + row;
    const float NEG = -RealTraits<float>::huge();
    uint4 ex = make_uint4(0u, 0u, 0u, 0u);       // this user's exclusion bits of the tile's 128 items
    if (MODE != MODE_SCORES && out.excl_bits && u < U)
        ex = __ldg(reinterpret_cast<const uint4*>(out.excl_bits + (size_t)u * out.excl_words + (j0 >> 5)));
    const float tau = (MODE == MODE_EMIT && u < U) ? out.tau[u] : 0.f;
    for (int c = 0; c < TN / 32; c++) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), r);
        if (u >= U) continue;
        const size_t jc = j0 + c * 32;
        if (MODE == MODE_SCORES) {
            float* srow = out.scores + (size_t)u * n + jc;
            int* irow = out.ids + (size_t)u * n + jc;
#pragma unroll
            for (int i = 0; i < 32; i++)
                if (jc + i < n) { srow[i] = __uint_as_float(r[i]); irow[i] = (int)(jc + i); }
        } else {
            const uint32_t exw = c == 0 ? ex.x : (c == 1 ? ex.y : (c == 2 ? ex.z : ex.w));
            // valid = inside [0, n) and not excluded
            uint32_t valid = ~exw;
            if (jc + 32 > n) valid &= (jc >= n) ? 0u : ((1u << (unsigned)(n - jc)) - 1u);
            if (MODE == MODE_GROUPMAX) {
#pragma unroll
                for (int g = 0; g < 32 / GROUP; g++) {
                    float m = NEG;
#pragma unroll
                    for (int i = 0; i < GROUP; i++) {
                        const float v = __uint_as_float(r[g * GROUP + i]);
                        if ((valid >> (g * GROUP + i)) & 1u) m = fmaxf(m, v);
                    }
                    const size_t gi = jc / GROUP + g;
                    if (gi < out.ngroups) out.gmax[(size_t)u * out.ngroups + gi] = m;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const float v = __uint_as_float(r[i]);
                    if (((valid >> i) & 1u) && v >= tau) {
                        const int pos = atomicAdd(out.cand_cnt + u, 1);
                        if (pos < CAND_CAP) {
                            out.cand_sc[(size_t)u * CAND_CAP + pos] = v;
                            out.cand_id[(size_t)u * CAND_CAP + pos] = (int)(jc + i);
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TN) : "memory");
}

// Exact re-scoring + final ordering of the per-user candidates.  One CTA (128 threads) per user.
//   cand_ids / cand_approx : the first M entries of the user's approx-sorted row (stride n)
// The M candidates are re-scored with the reference's left-to-right non-contracted sum, ordered by
// (score desc, id asc) with a bitonic network in shared memory, and the first n_top written out.
// flag[u] = 1 when the TF32 error bound cannot exclude an item outside the candidates.
__global__ void __launch_bounds__(128) rescore_select_kernel(const float* __restrict__ Asel, const float* __restrict__ B,
                                                             int k, int ldf, const int* __restrict__ cand_ids,
                                                             const float* __restrict__ cand_approx, size_t n, int M,
                                                             int n_avail_is_M, int n_top, float rel_bound,
                                                             long long* __restrict__ out_ids,
                                                             float* __restrict__ out_scores, int* __restrict__ flag)
{
    __shared__ float ss[256];
    __shared__ int si[256];
    __shared__ float a[256];
    const int u = blockIdx.x, tid = threadIdx.x;
    for (int c = tid; c < k; c += blockDim.x) a[c] = Asel[(size_t)u * ldf + c];
    __syncthreads();
    const float NEG = -RealTraits<float>::huge();
    for (int c = tid; c < 256; c += blockDim.x) {
        float s = NEG; int id = 0x7fffffff;
        if (c < M) {
            const float ap = cand_approx[(size_t)u * n + c];
            if (ap > NEG) {
                id = cand_ids[(size_t)u * n + c];
                const float* b = B + (size_t)id * ldf;
                float acc = 0;
                for (int i = 0; i < k; i++) acc = add_rn(acc, mul_rn(a[i], b[i]));
                s = acc;
            }
        }
        ss[c] = s; si[c] = id;
    }
    __syncthreads();
    for (int len = 2; len <= 256; len <<= 1)
        for (int j = len >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < 128; t += blockDim.x) {
                const int lo = 2 * t - (t & (j - 1)), hi = lo + j;
                const bool desc_block = ((lo & len) == 0);
                const float s0 = ss[lo], s1 = ss[hi];
                const int i0 = si[lo], i1 = si[hi];
                const bool before = (s0 > s1) || (s0 == s1 && i0 < i1);   // lo already precedes hi
                if (before != desc_block) { ss[lo] = s1; ss[hi] = s0; si[lo] = i1; si[hi] = i0; }
            }
            __syncthreads();
        }
    for (int c = tid; c < n_top; c += blockDim.x) {
        out_ids[(size_t)u * n_top + c] = si[c];
        if (out_scores) out_scores[(size_t)u * n_top + c] = ss[c];
    }
    if (tid == 0) {
        int bad = 0;
        if (!n_avail_is_M) {
            // every item outside the candidates has TF32 score <= a_min, hence exact score <= a_min * (1 + rel_bound)
            const float a_min = cand_approx[(size_t)u * n + (M - 1)];
            const float e_n = ss[n_top - 1];
            if (!(a_min == NEG) && !(e_n > a_min * (1.0f + rel_bound))) bad = 1;
        }
        flag[u] = bad;
    }
}

// bits[u, j] = 1 for every item j in user (user0 + u)'s exclusion list.  grid = (x, users of the chunk)
template <class IX>
__global__ void exclusion_bitmap_kernel(uint32_t* __restrict__ bits, size_t words, const IX* __restrict__ excl_ptr,
                                        const IX* __restrict__ excl_ix, size_t user0, size_t n)
{
    const size_t u = blockIdx.y;
    const size_t beg = (size_t)excl_ptr[user0 + u], end = (size_t)excl_ptr[user0 + u + 1];
    for (size_t t = beg + blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < end; t += (size_t)gridDim.x * blockDim.x) {
        const size_t j = (size_t)excl_ix[t];
        if (j < n) atomicOr(bits + u * words + (j >> 5), 1u << (j & 31));
    }
}

// tau[u] = the M-th largest of the user's group maxima (most-significant-digit radix select on the
// order-preserving keys, 8 bits a round).  One CTA of 256 threads per user.
__global__ void __launch_bounds__(256) select_threshold_kernel(const float* __restrict__ gmax, size_t ngroups, int M,
                                                               float* __restrict__ tau)
{
    __shared__ unsigned int hist[256];
    __shared__ uint32_t s_prefix;
    __shared__ int s_want;
    const float* g = gmax + (size_t)blockIdx.x * ngroups;
    const int tid = threadIdx.x;
    if (tid == 0) { s_prefix = 0u; s_want = M; }
    uint32_t mask = 0u;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[tid] = 0u;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        for (size_t i = tid; i < ngroups; i += 256) {
            const uint32_t key = score_key(g[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {     // walk the digits from the top until `want` keys have been passed
            int want = s_want, d = 255;
            for (; d > 0; d--) {
                const int c = (int)hist[d];
                if (c >= want) break;
                want -= c;
            }
            s_prefix = prefix | ((uint32_t)d << shift);
            s_want = want;
        }
        mask |= 255u << shift;
        __syncthreads();
    }
    if (tid == 0) tau[blockIdx.x] = key_score(s_prefix);
}

// Order one user's candidates by (score desc, id asc) and keep the best CAND_TOP (padded with -huge).
// overflow[u] = 1 when more than CAND_CAP items reached the threshold (ties at tau): redo exactly.
__global__ void __launch_bounds__(512) sort_candidates_kernel(const float* __restrict__ cand_sc,
                                                              const int* __restrict__ cand_id,
                                                              const int* __restrict__ cand_cnt,
                                                              float* __restrict__ top_sc, int* __restrict__ top_id,
                                                              int* __restrict__ overflow)
{
    __shared__ float ss[CAND_CAP];
    __shared__ int si[CAND_CAP];
    const int u = blockIdx.x, tid = threadIdx.x;
    const int cnt = cand_cnt[u];
    const float NEG = -RealTraits<float>::huge();
    const int m = cnt < CAND_CAP ? cnt : CAND_CAP;
    int len_pad = CAND_TOP;                          // sort only as much as is filled (power of two)
    while (len_pad < m) len_pad <<= 1;
    for (int c = tid; c < len_pad; c += blockDim.x) {
        ss[c] = c < m ? cand_sc[(size_t)u * CAND_CAP + c] : NEG;
        si[c] = c < m ? cand_id[(size_t)u * CAND_CAP + c] : 0x7fffffff;
    }
    __syncthreads();
    for (int len = 2; len <= len_pad; len <<= 1)
        for (int j = len >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < len_pad / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (j - 1)), hi = lo + j;
                const bool desc_block = ((lo & len) == 0);
                const float s0 = ss[lo], s1 = ss[hi];
                const int i0 = si[lo], i1 = si[hi];
                const bool before = (s0 > s1) || (s0 == s1 && i0 < i1);
                if (before != desc_block) { ss[lo] = s1; ss[hi] = s0; si[lo] = i1; si[hi] = i0; }
            }
            __syncthreads();
        }
    for (int c = tid; c < CAND_TOP; c += blockDim.x) {
        top_sc[(size_t)u * CAND_TOP + c] = ss[c];
        top_id[(size_t)u * CAND_TOP + c] = si[c];
    }
    if (tid == 0) overflow[u] = cnt > CAND_CAP ? 1 : 0;
}

__global__ void any_negative_kernel(const float* __restrict__ x, size_t n, int* __restrict__ flag)
{
    int f = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        f |= (x[i] < 0.f);
    if (f) atomicOr(flag, 1);
}

}  // namespace tc
}  // namespace pmf
