// Tensor-core iALS++ row solve for sm_100a (d = 128, block_size 32): the Blackwell-native path of
// lib/algo_impl/als/als.cc:211-358.
//
// Algebra.  With M = G + reg*I + sum_c w_c q_c q_c^T (w = alpha*v) and b = sum_c w_c q_c, the reference's block
// right-hand side (als.cc:296,303-308) is  g_B = (x M)[B] - b[B]  for the CURRENT x (Yui_c == x.q_c at every point of
// its loop), its CG operator (als.cc:278,330-336) is M[B,B], and "p_blk -= x; Yui -= q_blk.x" (als.cc:346-350) keeps
// that invariant.  So one row is: form the explicit d x d matrix once, then run the block Gauss-Seidel sweep with the
// fixed 3-step CG on 32 x 32 diagonal blocks -- O(d^2) work per row that does not depend on the row length, and the
// per-nnz work collapses into one rank-1 update  M += (sqrt(w) q)(sqrt(w) q)^T, a dense contraction: tensor cores.
//
// Pipeline of one persistent CTA (one per SM, 14 warps, warp-specialised, mbarrier hand-offs only):
//   producer warp   : walks its rows, one 512-byte cp.async.bulk (TMA, SASS UBLKCP) per gathered opposite-factor
//                     row into a ring of raw fp32 tiles (32 rows), completion by mbarrier expect_tx;
//   4 convert warps : raw tile -> s = sqrt|w| q, split into a tf32 head and tail (3xTF32: hi.hi + hi.lo + lo.hi carries
//                     ~2^-22 relative error, i.e. fp32-grade), written as MN-major operand slabs in the only layout
//                     tcgen05 takes for MN-major tf32 (128-byte swizzle with 32-byte atomicity: rows of 32 columns, 4 k
//                     per atom); they also accumulate b = sum w q (exact fp32) and the loss pieces;
//   MMA warp        : one lane issues tcgen05.mma kind::tf32 (M = N = 128, K = 8; SASS UTCHMMA), accumulating the row's
//                     matrix in tensor memory; entries with negative weight travel in their own tiles and are
//                     subtracted with the instruction descriptor's negate-A bit;
//   2 x 4 epilogue warps : thread j owns matrix row j (tensor-memory lane j): tcgen05.ld (SASS LDTM) the accumulator and
//                     the resident G + reg*I (tensor memory columns 0..127), h = M x - b, then per 32-column block the
//                     3-step CG in the owning warp and a rank-32 update of the later blocks' h.  Three accumulators
//                     (3 x 128 columns) rotate, so the epilogue of rows i, i+1 overlaps the contraction of row i+2.
// Rows of any length stream through (no per-nnz state), which removes the long-row cliff of the SIMT classes; rows
// longer than the split threshold are cut into chunks whose partial matrices are summed in global memory and solved
// by als_explicit_solve_kernel (als_explicit.cuh).
#pragma once
#include "als_generic.cuh"
#include "bfl_common.cuh"
#include "sm100_ptx.cuh"

namespace bfl {
namespace tc {

using namespace sm100;

constexpr int WARPS = 14, THREADS = WARPS * 32;
constexpr int W_PROD = 0, W_MMA = 1, W_CONV = 2, N_CONV = 4, W_EPI = 6;   // warps 6..9 group 0, 10..13 group 1
constexpr int NR = 4;      // raw stages
constexpr int NO = 3;      // operand stages
constexpr int NBV = 8;     // ring of per-row vectors handed from the convert warps to the epilogue
// d = 128: 32 gathered rows per stage, three 128-column accumulators (fused solve) behind the resident G + reg I;
// d = 256 (split-row mode only): 16 rows per stage, one accumulator set of 384 columns: rows 0..127 x all 256 columns
// and rows 128..255 x columns 128..255 (the remaining quadrant is the transpose of the first one's right half)
template <int D>
struct Cfg {
    static constexpr int TILE = D == 128 ? 32 : 16;      // gathered rows per stage
    static constexpr int NACC = D == 128 ? 3 : 1;        // accumulator sets in tensor memory
    static constexpr int KG = TILE / 4;                   // groups of 4 gathered rows (one swizzle atom deep)
    static constexpr int JS = 16 / KG;                    // column-chunk selectors per (k, chunk parity)
    static constexpr int NPART = D == 128 ? 2 : 1;        // partial b vectors handed to the epilogue (summed there)
    static constexpr int KSTEP_FLOATS = 8 * D;            // one K = 8 slab of an operand array
};
constexpr int NACC_MAX = 3;
constexpr uint64_t SWZ_128B_BASE32B = 1ull << 61;   // matrix-descriptor layout type 1
constexpr uint32_t F_FIRST = 1u << 8, F_LAST = 1u << 9, F_NEG = 1u << 10, F_STOP = 1u << 11;

template <int D>
struct Smem {
    static constexpr int RAWP = D + 8;                 // floats; 32-byte pad: conflict-free 128-bit reads of the convert map
    static constexpr int TILE = Cfg<D>::TILE;
    alignas(1024) float op[NO][2][TILE * D];           // [hi|lo][k-step slab]: swizzled MN-major atoms (see op_offset)
    alignas(128) float raw[NR][TILE * RAWP];
    alignas(16) float bvec[NBV][Cfg<D>::NPART][D];     // b = sum w q (partials over the convert warps)
    alignas(16) float sumq[NBV][Cfg<D>::NPART][D];     // sum q (loss only)
    alignas(16) float xs[2][D];                        // 128-bit reads: every vector below is 16-byte aligned
    alignas(16) float pv[2][32];
    alignas(16) float dl[2][2][32];
    alignas(16) float sw[NR][TILE];                    // sign(w) sqrt|w| per slot
    float wsum[NBV][2];                                // sum w (loss only; partials)
    uint32_t meta_raw[NR];
    uint32_t meta_op[NO];
    int badf[2][4];
    alignas(8) uint64_t raw_full[NR], raw_empty[NR], op_full[NO], op_empty[NO], acc_full[NACC_MAX], acc_empty[NACC_MAX];
    uint32_t tmem_base;
};

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

// sums v[i] (32 values per lane) over the 16 lanes that share lane bit 2 (xor offsets 16, 8, 2, 1); afterwards a lane
// holds in v[0], v[1] the totals of indices (b4 << 4 | b3 << 3 | b1 << 2 | b0 << 1) + {0, 1}, b_x = bit x of the lane id
__device__ __forceinline__ void transpose_reduce16(float (&v)[32], int lane) {
    int nv = 32;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int off = s == 0 ? 16 : (s == 1 ? 8 : (s == 2 ? 2 : 1));
        nv >>= 1;
        const bool up = lane & off;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < nv) {
                const float send = up ? v[i] : v[i + nv];
                const float keep = up ? v[i + nv] : v[i];
                v[i] = keep + __shfl_xor_sync(FULL, send, off);
            }
        }
    }
}

// PARTIAL = false: rows of a.row_list[row_begin..row_end) are solved in place.
// PARTIAL = true : the list holds chunk items of long rows (pairs: row, chunk index); the chunk's matrix and vectors are
//                  added to scratch[slot] (slot = items[3*i+2]) and solved later by als_explicit_solve_kernel.
struct TcArgs {
    AlsArgs a;
    const int32_t* items;   // PARTIAL: triples (row, chunk, scratch slot)
    float* scratch;         // PARTIAL: per slot D*D matrix + D (b) + D (sum q) + 4 (sum w, ...) floats
    int64_t split;          // PARTIAL: chunk length in nnz
    int loss_axis1;         // compute_loss && axis == 1: also hand over sum q / sum w
};

template <int D>
__host__ __device__ constexpr size_t scratch_floats() { return (size_t)D * D + 2 * D + 4; }

template <int D, bool PARTIAL>
__global__ void __launch_bounds__(THREADS, 1) als_tc_kernel(TcArgs ta) {
    static_assert(D == 128 || (D == 256 && PARTIAL), "fused row solve: d = 128; split-row mode: d = 128 or 256");
    constexpr int TILE = Cfg<D>::TILE, NACC = Cfg<D>::NACC, KSTEP_FLOATS = Cfg<D>::KSTEP_FLOATS;
    extern __shared__ __align__(1024) unsigned char smem_raw_[];
    Smem<D>& S = *reinterpret_cast<Smem<D>*>(smem_raw_);
    const AlsArgs& a = ta.a;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(FULL, tid >> 5, 0);
    constexpr int RAWP = Smem<D>::RAWP;

    if (tid == 0) {
        for (int i = 0; i < NR; ++i) {
            mbar_init(&S.raw_full[i], 1);
            mbar_init(&S.raw_empty[i], N_CONV * 32);
        }
        for (int i = 0; i < NO; ++i) {
            mbar_init(&S.op_full[i], N_CONV * 32);
            mbar_init(&S.op_empty[i], 1);
        }
        for (int i = 0; i < NACC; ++i) {
            mbar_init(&S.acc_full[i], 1);
            mbar_init(&S.acc_empty[i], 128);
        }
        mbar_init_fence();
    }
    if (warp == W_MMA) tmem_alloc(&S.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    // G + reg I -> tensor memory columns [0, D) (epilogue group 0; thread j holds matrix row j)
    if (!PARTIAL && warp >= W_EPI && warp < W_EPI + 4) {
        const int q = warp & 3, j = q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.G + (size_t)j * D + c * 32 + i));
                v[i] = g4.x; v[i + 1] = g4.y; v[i + 2] = g4.z; v[i + 3] = g4.w;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += (c * 32 + i == j) ? a.reg : 0.f;
            tmem_st32(tmem + ((uint32_t)(q * 32) << 16) + c * 32, v);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int64_t nitems = a.row_end - a.row_begin;
    const int64_t my_first = (int64_t)blockIdx.x;
    const int64_t stride = gridDim.x;

    // item -> (row, first entry, count)
    auto item_info = [&](int64_t it, int& row, int64_t& beg, int64_t& n, int& slot) {
        if (PARTIAL) {
            const int32_t* p = ta.items + 3 * (a.row_begin + it);
            row = p[0];
            const int64_t rb = row == 0 ? 0 : a.indptr[row - 1];
            const int64_t rn = a.indptr[row] - rb;
            beg = rb + (int64_t)p[1] * ta.split;
            n = min(ta.split, rn - (int64_t)p[1] * ta.split);
            slot = p[2];
        } else {
            row = a.row_list[a.row_begin + it];
            beg = row == 0 ? 0 : a.indptr[row - 1];
            n = a.indptr[row] - beg;
            slot = 0;
        }
    };

    if (warp == W_PROD) {
        // ================= producer: gathers =================
        uint32_t rs = 0, rph = 0;   // stage, phase of raw_empty
        auto emit = [&](unsigned mask, uint32_t flags, int32_t key, float w) {
            const int cnt = __popc(mask);
            const bool mine = (mask >> lane) & 1u;
            const int slot = __popc(mask & ((1u << lane) - 1u));
            mbar_wait(&S.raw_empty[rs], rph ^ 1u);
            if (mine) S.sw[rs][slot] = copysignf(sqrtf(fabsf(w)), w);
            __syncwarp();
            if (lane == 0) {
                S.meta_raw[rs] = (uint32_t)cnt | flags;
                mbar_arrive_expect_tx(&S.raw_full[rs], (uint32_t)cnt * D * 4);
            }
            __syncwarp();
            if (mine) bulk_g2s(&S.raw[rs][slot * RAWP], a.Y + (int64_t)key * a.ld, D * 4, &S.raw_full[rs]);
            if (++rs == NR) { rs = 0; rph ^= 1u; }
        };
        int row, slot_unused;
        int64_t beg = 0, n = 0;
        int nrow = 0;
        int64_t nbeg = 0, nn = 0;
        if (my_first < nitems) item_info(my_first, row, beg, n, slot_unused);
        for (int64_t it = my_first; it < nitems; it += stride) {
            if (it + stride < nitems) item_info(it + stride, nrow, nbeg, nn, slot_unused);   // look-ahead
            // first chunk's entries
            const bool inl = lane < TILE;
            int64_t idx = lane;
            int32_t key = (inl && idx < n) ? a.keys[beg - a.shift + idx] : 0;
            float w = (inl && idx < n) ? a.vals[beg - a.shift + idx] * a.alpha : 0.f;
            for (int64_t c0 = 0; c0 < n; c0 += TILE) {
                // prefetch the next chunk of this row
                const int64_t idx2 = c0 + TILE + lane;
                const int32_t key2 = (inl && idx2 < n) ? a.keys[beg - a.shift + idx2] : 0;
                const float w2 = (inl && idx2 < n) ? a.vals[beg - a.shift + idx2] * a.alpha : 0.f;
                const bool valid = inl && c0 + lane < n;
                const unsigned pm = __ballot_sync(FULL, valid && !(w < 0.f));
                const unsigned nm = __ballot_sync(FULL, valid && (w < 0.f));
                const bool lastc = c0 + TILE >= n;
                uint32_t fl = (c0 == 0 ? F_FIRST : 0u);
                if (pm) {
                    emit(pm, fl | ((lastc && !nm) ? F_LAST : 0u), key, w);
                    fl = 0u;
                }
                if (nm) emit(nm, fl | F_NEG | (lastc ? F_LAST : 0u), key, w);
                key = key2;
                w = w2;
            }
            row = nrow; beg = nbeg; n = nn;
        }
        // stop marker
        mbar_wait(&S.raw_empty[rs], rph ^ 1u);
        if (lane == 0) {
            S.meta_raw[rs] = F_STOP;
            mbar_arrive(&S.raw_full[rs]);
        }
    } else if (warp == W_MMA) {
        // ================= MMA issue =================
        uint32_t os = 0, oph = 0, acc = 0, aph = 0;
        const uint32_t idesc = idesc_tf32_mn(128, 128), idesc_neg = idesc | (1u << 13);
        const uint32_t idesc_w = idesc_tf32_mn(128, 256), idesc_w_neg = idesc_w | (1u << 13);   // d = 256: N = 256
        bool row_open = false;
        for (;;) {
            mbar_wait(&S.op_full[os], oph);
            tc_fence_after();
            const uint32_t meta = S.meta_op[os];
            if (meta & F_STOP) break;
            if (meta & F_FIRST) {
                mbar_wait(&S.acc_empty[acc], aph ^ 1u);
                tc_fence_after();
                row_open = false;
            }
            if (lane == 0) {
                const int ksteps = (int)(meta & 0xffu);
                const uint32_t id = (meta & F_NEG) ? idesc_neg : idesc;
                const uint32_t hi = s32(&S.op[os][0][0]), lo = s32(&S.op[os][1][0]);
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t acc0 = (row_open || ks > 0) ? 1u : 0u;
                    // MN-major tf32 operand: SWIZZLE_128B_BASE32B atoms of 32 (M/N) x 4 (K) values = 512 B; the next 4 k
                    // 512 B further (stride-dimension offset), the next 32 columns 1024 B further (leading-dimension offset)
                    const uint64_t dh = smem_desc(hi + ks * KSTEP_FLOATS * 4, 1024, 512) | SWZ_128B_BASE32B;
                    const uint64_t dl = smem_desc(lo + ks * KSTEP_FLOATS * 4, 1024, 512) | SWZ_128B_BASE32B;
                    if (D == 128) {
                        const uint32_t dcol = tmem + D * (1 + acc);
                        mma_tf32(dcol, dh, dh, id, acc0);
                        mma_tf32(dcol, dh, dl, id, 1u);
                        mma_tf32(dcol, dl, dh, id, 1u);
                    } else {
                        // rows 0..127 x columns 0..255 -> tensor-memory columns [0, 256)
                        const uint32_t idw = (meta & F_NEG) ? idesc_w_neg : idesc_w;
                        mma_tf32(tmem, dh, dh, idw, acc0);
                        mma_tf32(tmem, dh, dl, idw, 1u);
                        mma_tf32(tmem, dl, dh, idw, 1u);
                        // rows 128..255 x columns 128..255 -> tensor-memory columns [256, 384): the slab's second half
                        const uint64_t eh = smem_desc(hi + ks * KSTEP_FLOATS * 4 + 4096, 1024, 512) | SWZ_128B_BASE32B;
                        const uint64_t el = smem_desc(lo + ks * KSTEP_FLOATS * 4 + 4096, 1024, 512) | SWZ_128B_BASE32B;
                        mma_tf32(tmem + 256, eh, eh, id, acc0);
                        mma_tf32(tmem + 256, eh, el, id, 1u);
                        mma_tf32(tmem + 256, el, eh, id, 1u);
                    }
                }
                mma_commit(&S.op_empty[os]);
                if (meta & F_LAST) mma_commit(&S.acc_full[acc]);
            }
            __syncwarp();
            row_open = true;
            if (meta & F_LAST) {
                if (++acc == NACC) { acc = 0; aph ^= 1u; }
            }
            if (++os == NO) { os = 0; oph ^= 1u; }
        }
    } else if (warp >= W_CONV && warp < W_CONV + N_CONV) {
        // ================= convert: raw fp32 -> scaled tf32 head/tail operand slabs =================
        // thread -> gathered row k = 4*kg + k4 of the tile and the 16-byte column chunks j = jpar + 2*(jsel + JS*i).
        // lane bits 0-1 = k % 4 and bit 2 = chunk parity make every quarter-warp hit 8 distinct 16-byte bank groups,
        // both on the raw read (pitch D+8 floats) and on the swizzled operand write.
        const int ct = tid - W_CONV * 32;
        constexpr int KG = Cfg<D>::KG, JS = Cfg<D>::JS, NPART = Cfg<D>::NPART;
        const int k4 = ct & 3, jpar = (ct >> 2) & 1, rest = ct >> 3;
        const int kg = rest % KG, jsel = rest / KG;
        const int k = 4 * kg + k4;
        const int part = NPART == 2 ? ((ct >> 5) & 1) : 0;          // which half of the k groups this warp covers
        // element (k, column m = 4j..4j+3) of a K = 8 slab, in floats:
        //   (m/32)*256 + ((k%8)/4)*128 + (k%4)*32 + ((((m%32)/8) ^ (k%4))*8) + m%8
        const int op_base = (k >> 3) * KSTEP_FLOATS + ((k >> 2) & 1) * 128 + k4 * 32 + jpar * 4;
        uint32_t rs = 0, rph = 0, os = 0, oph = 0, bslot = 0;
        float bacc[32], qacc[32];
        float wacc = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) { bacc[i] = 0.f; qacc[i] = 0.f; }
        for (;;) {
            mbar_wait(&S.raw_full[rs], rph);
            const uint32_t meta = S.meta_raw[rs];
            mbar_wait(&S.op_empty[os], oph ^ 1u);
            if (meta & F_STOP) {
                if (ct == 0) S.meta_op[os] = F_STOP;
                mbar_arrive(&S.op_full[os]);
                break;
            }
            const int cnt = (int)(meta & 0xffu), ksteps = (cnt + 7) >> 3;
            if (meta & F_FIRST) {
#pragma unroll
                for (int i = 0; i < 32; ++i) { bacc[i] = 0.f; qacc[i] = 0.f; }
                wacc = 0.f;
            }
            if (k < ksteps * 8) {
                const bool valid = k < cnt;
                const float swv = valid ? S.sw[rs][k] : 0.f;
                const float sa = fabsf(swv);
                if (jpar == 0 && jsel == 0) wacc = fmaf(swv, sa, wacc);
                float* hi = &S.op[os][0][op_base];
                float* lo = &S.op[os][1][op_base];
                const float* rp = &S.raw[rs][k * RAWP];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = jpar + 2 * (jsel + JS * i);   // 16-byte column chunk
                    const int o = (j >> 3) * 256 + ((((j & 7) >> 1) ^ k4) * 8);
                    float4 q = valid ? *reinterpret_cast<const float4*>(rp + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 s = make_float4(q.x * sa, q.y * sa, q.z * sa, q.w * sa), h, l;
                    split_tf32(s.x, h.x, l.x);
                    split_tf32(s.y, h.y, l.y);
                    split_tf32(s.z, h.z, l.z);
                    split_tf32(s.w, h.w, l.w);
                    *reinterpret_cast<float4*>(hi + o) = h;
                    *reinterpret_cast<float4*>(lo + o) = l;
                    bacc[4 * i + 0] = fmaf(swv, s.x, bacc[4 * i + 0]);
                    bacc[4 * i + 1] = fmaf(swv, s.y, bacc[4 * i + 1]);
                    bacc[4 * i + 2] = fmaf(swv, s.z, bacc[4 * i + 2]);
                    bacc[4 * i + 3] = fmaf(swv, s.w, bacc[4 * i + 3]);
                    if (ta.loss_axis1) {
                        qacc[4 * i + 0] += q.x; qacc[4 * i + 1] += q.y; qacc[4 * i + 2] += q.z; qacc[4 * i + 3] += q.w;
                    }
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(&S.raw_empty[rs]);
            if (meta & F_LAST) {
                // value index 4*i + comp <-> column 4*j(i) + comp; the 16 lanes sharing (jpar, jsel) are reduced
                transpose_reduce16(bacc, lane);
                if (ta.loss_axis1) transpose_reduce16(qacc, lane);
                const int vi0 = ((lane >> 4) & 1) << 4 | ((lane >> 3) & 1) << 3 | ((lane >> 1) & 1) << 2 | (lane & 1) << 1;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int vi = vi0 + e;
                    const int col = 4 * (jpar + 2 * (jsel + JS * (vi >> 2))) + (vi & 3);
                    S.bvec[bslot][part][col] = bacc[e];
                    if (ta.loss_axis1) S.sumq[bslot][part][col] = qacc[e];
                }
                if (ta.loss_axis1 && jsel == 0) {   // warps whose threads (jpar == 0) see each gathered row once
                    const float ws = warp_sum(jpar == 0 ? wacc : 0.f);
                    if (lane == 0) S.wsum[bslot][part] = ws;
                }
                bslot = (bslot + 1) & (NBV - 1);
            }
            if (ct == 0) S.meta_op[os] = (uint32_t)ksteps | (meta & (F_FIRST | F_LAST | F_NEG));
            mbar_arrive(&S.op_full[os]);
            if (++rs == NR) { rs = 0; rph ^= 1u; }
            if (++os == NO) { os = 0; oph ^= 1u; }
        }
    } else {
        // ================= epilogue: explicit-matrix block Gauss-Seidel / CG =================
        const int g = (warp - W_EPI) >> 2;      // group
        const int q = warp & 3;                  // tensor-memory lane quarter == column block owned by this warp
        const int j = q * 32 + lane;             // matrix row
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float* xs = S.xs[g];
        float* pv = S.pv[g];
        double l_nume = 0.0, l_deno = 0.0;
        const float tol = a.tol;
        int row = 0, slot = 0, nrow = 0, nslot = 0;
        int64_t beg, n = 0, nbeg, nn = 0;
        int64_t seq = g;
        float xj = 0.f, nxj = 0.f;
        if (my_first + seq * stride < nitems) {
            item_info(my_first + seq * stride, row, beg, n, slot);
            if (!PARTIAL) xj = a.X[(int64_t)row * a.ld + j];
        }
        for (; my_first + seq * stride < nitems; seq += 2) {
            const int64_t nit = my_first + (seq + 2) * stride;
            if (nit < nitems) {
                item_info(nit, nrow, nbeg, nn, nslot);
                if (!PARTIAL) nxj = a.X[(int64_t)nrow * a.ld + j];
            }
            const uint32_t acc = (uint32_t)(seq % NACC), aph = (uint32_t)((seq / NACC) & 1);
            const uint32_t bs = (uint32_t)(seq & (NBV - 1));
            mbar_wait(&S.acc_full[acc], aph);
            tc_fence_after();
            if constexpr (PARTIAL) {
                // add this chunk's matrix / vectors to the row's scratch block (float atomics: the chunks of one row
                // are summed in arrival order)
                float* sc = ta.scratch + (size_t)slot * scratch_floats<D>();
                if constexpr (D == 128) {
                    const uint32_t dbase = tmem + lane_off + D * (1 + acc);
#pragma unroll 1
                    for (int c = 0; c < D / 32; ++c) {
                        float v[32];
                        tmem_ld32(dbase + c * 32, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) atomicAdd(sc + (size_t)j * D + c * 32 + i, v[i]);
                    }
                } else {
                    // tensor-memory columns [0,256): matrix rows 0..127; [256,384): rows 128..255 x columns 128..255;
                    // rows 128..255 x columns 0..127 are the transpose of rows 0..127 x columns 128..255
#pragma unroll 1
                    for (int c = 0; c < 8; ++c) {
                        float v[32];
                        tmem_ld32(tmem + lane_off + c * 32, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            atomicAdd(sc + (size_t)j * D + c * 32 + i, v[i]);
                            if (c >= 4) atomicAdd(sc + (size_t)(c * 32 + i) * D + j, v[i]);
                        }
                    }
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        float v[32];
                        tmem_ld32(tmem + lane_off + 256 + c * 32, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) atomicAdd(sc + (size_t)(128 + j) * D + 128 + c * 32 + i, v[i]);
                    }
                }
                for (int jj = j; jj < D; jj += 128) {
                    float bb = 0.f, qq = 0.f;
#pragma unroll
                    for (int pp = 0; pp < Cfg<D>::NPART; ++pp) {
                        bb += S.bvec[bs][pp][jj];
                        if (ta.loss_axis1) qq += S.sumq[bs][pp][jj];
                    }
                    atomicAdd(sc + (size_t)D * D + jj, bb);
                    if (ta.loss_axis1) atomicAdd(sc + (size_t)D * D + D + jj, qq);
                }
                if (ta.loss_axis1 && j == 0)
                    atomicAdd(sc + (size_t)D * D + 2 * D, S.wsum[bs][0] + (Cfg<D>::NPART == 2 ? S.wsum[bs][1] : 0.f));
                tc_fence_before();
                mbar_arrive(&S.acc_empty[acc]);
                row = nrow; slot = nslot; n = nn;
                continue;
            }
            const uint32_t dbase = tmem + lane_off + 128 * (1 + acc);
            xs[j] = xj;
            float bj = S.bvec[bs][0][j];
            if (Cfg<D>::NPART == 2) bj += S.bvec[bs][Cfg<D>::NPART - 1][j];
            group_sync(g);
            // ---- h = (G + reg I) x + D x - b; keep the diagonal block of M in registers ----
            float hG = 0.f, hD = 0.f;
            float md[32];
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {
                float dv[32], gv[32];
                tmem_ld32(dbase + c * 32, dv);
                tmem_ld32(tmem + lane_off + c * 32, gv);
                tmem_wait_ld();
                float h0 = 0.f, h1 = 0.f, g0 = 0.f, g1 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 x4 = *reinterpret_cast<const float4*>(xs + c * 32 + i);
                    h0 = fmaf(dv[i], x4.x, h0); h1 = fmaf(dv[i + 1], x4.y, h1);
                    h0 = fmaf(dv[i + 2], x4.z, h0); h1 = fmaf(dv[i + 3], x4.w, h1);
                    g0 = fmaf(gv[i], x4.x, g0); g1 = fmaf(gv[i + 1], x4.y, g1);
                    g0 = fmaf(gv[i + 2], x4.z, g0); g1 = fmaf(gv[i + 3], x4.w, g1);
                }
                hD += h0 + h1;
                hG += g0 + g1;
                if (c == q) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) md[i] = dv[i] + gv[i];
                }
            }
            if (a.compute_loss) {
                // als.cc:298-321 with the pre-update row: reg*kappa*|x|^2 (both axes); item side additionally
                // x G x + sum_obs[(1+w)(yhat-1)^2 - yhat^2] = x G x + x D x - 2 x.(b + sum q) + (n + sum w)
                const float kappa = a.adaptive_reg ? (float)n : 1.0f;
                double t = (double)(kappa * a.reg * xj * xj);
                if (a.axis == 1) {
                    t += (double)xj * (double)(hG - a.reg * xj) + (double)xj * (double)hD -
                         2.0 * (double)xj * ((double)bj + (double)S.sumq[bs][0][j] +
                                             (Cfg<D>::NPART == 2 ? (double)S.sumq[bs][Cfg<D>::NPART - 1][j] : 0.0));
                    if (j == 0) {
                        const double ws = (double)S.wsum[bs][0] + (Cfg<D>::NPART == 2 ? (double)S.wsum[bs][1] : 0.0);
                        t += (double)n + ws;
                        l_deno += (double)a.Y_rows + ws;
                    }
                }
                l_nume += t;
            }
            float h = hG + hD - bj;
            // ---- block sweep ----
#pragma unroll 1
            for (int B = 0; B < D / 32; ++B) {
                if (q == B) {
                    float r = h, p = h, xv = 0.f;
                    float rsold = warp_sum(r * r);
                    bool act = rsold > tol;            // als.cc:329
#pragma unroll 1
                    for (int step = 0; step < 3; ++step) {
                        pv[lane] = p;
                        __syncwarp();
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 p4 = *reinterpret_cast<const float4*>(pv + i);
                            a0 = fmaf(md[i], p4.x, a0); a1 = fmaf(md[i + 1], p4.y, a1);
                            a2 = fmaf(md[i + 2], p4.z, a2); a3 = fmaf(md[i + 3], p4.w, a3);
                        }
                        __syncwarp();
                        const float Ap = (a0 + a1) + (a2 + a3);
                        const float pAp = warp_sum(p * Ap);
                        const float ss = act ? __fdividef(rsold, pAp) : 0.f;   // als.cc:337 (no eps)
                        xv = fmaf(ss, p, xv);
                        r = fmaf(-ss, Ap, r);
                        const float rsnew = warp_sum(r * r);
                        act = act && !(rsnew < tol);                            // als.cc:341
                        if (act) p = fmaf(__fdividef(rsnew, rsold), p, r);
                        rsold = act ? rsnew : rsold;
                    }
                    S.dl[g][B & 1][lane] = xv;
                    xs[j] -= xv;                        // als.cc:346
                }
                group_sync(g);
                if (q > B) {   // later blocks: h -= M[j, B] . delta
                    float dv[32], gv[32];
                    tmem_ld32(dbase + B * 32, dv);
                    tmem_ld32(tmem + lane_off + B * 32, gv);
                    tmem_wait_ld();
                    float u0 = 0.f, u1 = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 d4 = *reinterpret_cast<const float4*>(&S.dl[g][B & 1][i]);
                        u0 = fmaf(dv[i] + gv[i], d4.x, u0); u1 = fmaf(dv[i + 1] + gv[i + 1], d4.y, u1);
                        u0 = fmaf(dv[i + 2] + gv[i + 2], d4.z, u0); u1 = fmaf(dv[i + 3] + gv[i + 3], d4.w, u1);
                    }
                    h -= u0 + u1;
                }
            }
            // the accumulator is free again
            tc_fence_before();
            mbar_arrive(&S.acc_empty[acc]);
            // NaN/Inf guard (cf. als.cu:116-120), write the row (and the peers' replicas, fused exchange)
            float v = xs[j];
            const bool badw = __any_sync(FULL, !isfinite(v));
            if (lane == 0) S.badf[g][q] = badw;
            group_sync(g);
            const bool bad = S.badf[g][0] | S.badf[g][1] | S.badf[g][2] | S.badf[g][3];
            v = bad ? 0.f : v;
            a.X[(int64_t)row * a.ld + j] = v;
            for (int pr = 0; pr < a.n_peer; ++pr) a.peerX[pr][(int64_t)row * a.ld + j] = v;
            row = nrow; slot = nslot; n = nn; xj = nxj;
        }
        if (!PARTIAL && a.loss && a.compute_loss) {
            l_nume = warp_sum_d(l_nume);
            l_deno = warp_sum_d(l_deno);
            if (lane == 0 && (l_nume != 0.0 || l_deno != 0.0)) {
                atomicAdd(a.loss, l_nume);
                atomicAdd(a.loss + 1, l_deno);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tmem, 512);
}

// ---- host side ---------------------------------------------------------------------------------
inline bool tc_applicable(int optimizer_code, int d, int vdim, int block_size) {
    return optimizer_code == 8 && d == 128 && vdim == 128 && block_size == 32;
}
// split-row mode (rows beyond the SIMT kernels' cap): d = 128 and d = 256
inline bool tc_split_applicable(int optimizer_code, int d, int vdim, int block_size) {
    return optimizer_code == 8 && (d == 128 || d == 256) && vdim == d && block_size == 32;
}

// split-row items: every row of list[0..nrows) is cut into chunks of `split` entries -> triples (row, chunk, slot)
__global__ void tc_count_items_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ list, int64_t nrows,
                                      int64_t split, unsigned long long* total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += (int64_t)gridDim.x * blockDim.x) {
        const int row = list[i];
        const int64_t n = indptr[row] - (row == 0 ? 0 : indptr[row - 1]);
        atomicAdd(total, (unsigned long long)((n + split - 1) / split));
    }
}
__global__ void tc_fill_items_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ list, int64_t nrows,
                                     int64_t split, unsigned long long* cursor, int32_t* __restrict__ items) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += (int64_t)gridDim.x * blockDim.x) {
        const int row = list[i];
        const int64_t n = indptr[row] - (row == 0 ? 0 : indptr[row - 1]);
        const int64_t nc = (n + split - 1) / split;
        const unsigned long long pos = atomicAdd(cursor, (unsigned long long)nc);
        for (int64_t c = 0; c < nc; ++c) {
            items[3 * (pos + c) + 0] = row;
            items[3 * (pos + c) + 1] = (int32_t)c;
            items[3 * (pos + c) + 2] = (int32_t)i;
        }
    }
}

// accumulates the chunk matrices of the split rows into scratch (zeroed here); a.row_begin/row_end index `items`
template <int D>
int tc_launch_partial(const AlsArgs& a, const int32_t* items, int64_t nitems, float* scratch, int64_t nslots,
                      int64_t split, int num_sms, cudaStream_t st) {
    if (nitems <= 0) return BFL_OK;
    TcArgs ta;
    ta.a = a;
    ta.a.row_begin = 0;
    ta.a.row_end = nitems;
    ta.items = items;
    ta.scratch = scratch;
    ta.split = split;
    ta.loss_axis1 = (a.compute_loss && a.axis == 1) ? 1 : 0;
    BFL_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * scratch_floats<D>() * (size_t)nslots, st));
    const size_t smem = sizeof(Smem<D>);
    BFL_CUDA(cudaFuncSetAttribute(als_tc_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>(nitems, (int64_t)num_sms);
    als_tc_kernel<D, true><<<grid, THREADS, smem, st>>>(ta);
    BFL_LAUNCHED();
    return BFL_OK;
}

// solves the rows a.row_list[a.row_begin .. a.row_end) (any length > 0) with the fused tensor-core kernel
inline int tc_launch(const AlsArgs& a, int num_sms, cudaStream_t st) {
    const int64_t nrows = a.row_end - a.row_begin;
    if (nrows <= 0) return BFL_OK;
    TcArgs ta;
    ta.a = a;
    ta.items = nullptr;
    ta.scratch = nullptr;
    ta.split = 0;
    ta.loss_axis1 = (a.compute_loss && a.axis == 1) ? 1 : 0;
    const size_t smem = sizeof(Smem<128>);
    BFL_CUDA(cudaFuncSetAttribute(als_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>(nrows, (int64_t)num_sms);
    als_tc_kernel<128, false><<<grid, THREADS, smem, st>>>(ta);
    BFL_LAUNCHED();
    return BFL_OK;
}

}  // namespace tc
}  // namespace bfl
