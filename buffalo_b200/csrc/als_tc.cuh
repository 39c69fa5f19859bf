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
// Precision.  The contraction runs on tcgen05.mma kind::f16 with a two-term fp16 split of the scaled operand
// t = 2^e sqrt|w| q = head + tail (head: 11 significant bits, tail: the next 11): head.head + head.tail + tail.head
// leaves a relative error of ~2^-21 per product, fp32-grade like the reference's arithmetic, at twice the tensor rate and
// half the shared-memory traffic per entry of the equivalent 3xTF32 scheme (K = 16 per instruction instead of 8).  The
// power of two 2^e (tc_scale_kernel: from max|Y| and max|w| of the launch) places the largest operand just below 2^15, so
// nothing overflows fp16 and entries down to 2^-20 of the largest keep a normal tail; the accumulator is multiplied by
// 2^-2e (exact) when it is read.  b and the loss pieces are accumulated in plain fp32 from the unscaled rows.
//
// Pipeline of one persistent CTA (one per SM, 14 warps x 128 registers, warp-specialised, mbarrier hand-offs only, one
// elected arrival per warp; hand-offs are per GROUP of two consecutive tiles of the CTA's tile stream):
//   planner warp     : walks the CTA's rows; every level of the dependent load chain (row id -> offsets -> keys / values)
//                     is issued whole rows ahead of its use; per tile of 32 entries it writes the gather plan (keys,
//                     2^e sqrt|w|, w, count / flags) into a 16-slot ring -- it runs up to 8 groups ahead;
//   8 convert warps  : (a) gathers: two hand-offs ahead of its use, warp cw copies the rows cw, cw + 8, ... of a planned
//                     tile with one coalesced 16-byte-per-lane cp.async per 512 bytes (SASS LDGSTS) into a ring of raw fp32
//                     tiles (commit / wait groups + one mbarrier arrival per warp).  The first version gathered with
//                     512-byte cp.async.bulk copies (TMA): a divergent-address bulk copy compiles to an ELECT / R2UR / UBLKCP
//                     loop over the lanes, ~63-100 cycles per copy and warp -- one warp sustains 2.3 TB/s chip-wide, four
//                     or more the 7.0 TB/s HBM read ceiling (benchmarks/gather_probe.cu) -- so it needed six warps that
//                     did nothing else; spread over the convert warps the cp.async issue is a few instructions per tile;
//                     (b) convert: thread = (feature m, entry half): reads column m of the raw tile (conflict-free),
//                     scales, splits, and writes 16-byte groups of 8 consecutive k into the K-major un-swizzled operand
//                     slabs (8 x 16 B core matrices, conflict-free); accumulates b_m = sum w q_m (exact fp32) and the loss
//                     pieces in registers; groups of two full tiles take a branch-free straight-line path;
//   MMA warp         : executes the generic -> async proxy fence for the group it has acquired, then one lane issues
//                     tcgen05.mma kind::f16 (M = N = 128, K = 16; SASS UTCHMMA), accumulating the row's matrix in tensor
//                     memory; entries with negative weight travel in their own tiles and are subtracted with the
//                     instruction descriptor's negate-A bit;
//   4 epilogue warps : a systolic pipeline over rows.  Thread j owns matrix row j (tensor-memory lane j), warp q column
//                     block q: tcgen05.ld (SASS LDTM) the accumulator and the resident G + reg*I (tensor memory columns
//                     0..127), h = M x - b; then warp q folds the deltas of blocks 0..q-1 into its h as they are published,
//                     runs the 3-step CG of its own 32 x 32 block and publishes its delta.  Nothing in a row's sweep is
//                     a group-wide barrier, so warp 0 is already on the next row's block 0 while warp 3 finishes this
//                     one: up to two of the three accumulators are being drained while the third is being filled.
// Rows of any length stream through (no per-nnz state), which removes the long-row cliff of the SIMT classes; rows
// longer than the split threshold are cut into chunks whose partial matrices are summed in global memory and solved
// by als_explicit_solve_kernel (als_explicit.cuh).
#pragma once
#include "als_generic.cuh"
#include "bfl_common.cuh"
#include <type_traits>

#include "sm100_ptx.cuh"

namespace bfl {
namespace tc {

using namespace sm100;

// 14 warps: epilogue 0..3 (warp q owns tensor-memory lane quarter q), planner 4, MMA issue 5, convert 6..13 (two sets of four:
// each set takes half of a tile's entries)
constexpr int N_CONV = 8, KH = N_CONV / 4;
constexpr int W_EPI = 0, W_PLAN = 4, W_MMA = 5, W_CONV = 6;
constexpr int WARPS = W_CONV + N_CONV, THREADS = WARPS * 32;
constexpr int NXS = 8;     // x buffers of the epilogue pipeline (a fast warp publishes row r + 1 while a slow one reads row r - 3)
// Hand-offs are per GROUP of PT consecutive tiles of the CTA's tile stream (a group may span rows: every tile carries its own
// flags): with all the math switched off the barrier round trips of a one-tile hand-off still cost ~700 cycles per tile.
constexpr int PT = 2;      // tiles per hand-off group
constexpr int NR = 3;      // raw stages (groups)
constexpr int NO = 2;      // operand stages (groups)
constexpr int GATHER_AHEAD = 2;   // groups between a convert warp's gather issue and its use of the group (< NR)
constexpr int NPG = 8, NP = NPG * PT;   // gather-plan ring (small slots): the planner runs up to NPG groups ahead
constexpr int NBV = 8;     // ring of per-row vectors handed from the convert warps to the epilogue
// d = 128: 32 gathered rows per stage, three 128-column accumulators (fused solve) behind the resident G + reg I;
// d = 256 (split-row mode only): 16 rows per stage, one accumulator set of 384 columns: rows 0..127 x all 256 columns
// and rows 128..255 x columns 128..255 (the remaining quadrant is the transpose of the first one's right half)
template <int D>
struct Cfg {
    static constexpr int TILE = D == 128 ? 32 : 16;      // gathered rows per stage
    static constexpr int NACC = D == 128 ? 3 : 1;        // accumulator sets in tensor memory
    static constexpr int NF = D / 128;                    // features per convert thread
    static constexpr int LBO = D * 16;                    // bytes between the 8-k chunks of an operand slab
    static constexpr int OP_BYTES = TILE * D * 2;         // head (or tail) slab of one stage
};
constexpr int NACC_MAX = 3;
constexpr uint32_t F_FIRST = 1u << 8, F_LAST = 1u << 9, F_NEG = 1u << 10, F_STOP = 1u << 11;

template <int D>
struct Smem {
    static constexpr int TILE = Cfg<D>::TILE;
    // operand slab: element (feature m, entry k) at byte (k/8)*LBO + (m/8)*128 + (m%8)*16 + (k%8)*2
    alignas(1024) unsigned char op[NO][PT][2][Cfg<D>::OP_BYTES];   // [stage][tile of the group][head|tail]
    alignas(128) float raw[NR][PT][TILE * D];          // gathered rows, pitch D
    alignas(16) float bvec[NBV][KH][D];                // b = sum w q (one partial per convert set)
    alignas(16) float sumq[NBV][KH][D];                // sum q (loss only)
    alignas(16) float xs[NXS][D];                      // 128-bit reads: every vector below is 16-byte aligned
    alignas(16) float pv[4][32];                       // CG direction of the warp that owns the block
    alignas(16) float dl[NACC_MAX][4][32];             // [row slot][block]: the block's solution delta
    alignas(16) float sws[NP][TILE];                   // gather plan: 2^e sqrt|w| per slot
    alignas(16) float wv[NP][TILE];                    //              w per slot
    alignas(16) int32_t keys[NP][TILE];                //              gathered row per slot
    float wsum[NBV][KH];                               // sum w (loss only)
    uint32_t meta_raw[NP];                             //              count | flags
    uint32_t meta_op[NO][PT];
    int badrow[NXS];
    alignas(8) uint64_t plan_full[NPG], plan_empty[NPG], raw_full[NR], raw_empty[NR], op_full[NO], op_empty[NO];
    alignas(8) uint64_t acc_full[NACC_MAX], acc_empty[NACC_MAX], x_full[NXS], d_full[NACC_MAX][4];
    uint32_t tmem_base;
};

// PARTIAL = false: rows of a.row_list[row_begin..row_end) are solved in place.
// PARTIAL = true : the list holds chunk items of long rows (pairs: row, chunk index); the chunk's matrix and vectors are
//                  added to scratch[slot] (slot = items[3*i+2]) and solved later by als_explicit_solve_kernel.
struct TcArgs {
    AlsArgs a;
    const int32_t* items;   // PARTIAL: triples (row, chunk, scratch slot)
    float* scratch;         // PARTIAL: per slot D*D matrix + D (b) + D (sum q) + 4 (sum w, ...) floats
    int64_t split;          // PARTIAL: chunk length in nnz
    int debug;              // BFL_TC_DEBUG (timing experiments only; results are wrong): 1 no gathers, 2 no MMAs, 4 no convert math, 8 no epilogue math, 16 planner only
};

template <int D>
__host__ __device__ constexpr size_t scratch_floats() { return (size_t)D * D + 2 * D + 4; }

// operand scaling of one launch: out[0] = 2^e, out[1] = 2^-2e with 2^e max|Y| sqrt(max|w|) in [2^14, 2^15)
__global__ void tc_absmax_kernel(const float* __restrict__ p, size_t n, unsigned int* __restrict__ out) {
    float m = 0.f;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    const size_t lead = min(n, (size_t)(((16 - ((uintptr_t)p & 15)) & 15) / 4));   // scalars up to 16-byte alignment
    const size_t n4 = (n - lead) / 4;
    const float4* p4 = reinterpret_cast<const float4*>(p + lead);
    for (size_t i = gtid; i < n4; i += gsz) {
        const float4 v = __ldg(p4 + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (size_t i = gtid; i < lead; i += gsz) m = fmaxf(m, fabsf(p[i]));
    for (size_t i = lead + n4 * 4 + gtid; i < n; i += gsz) m = fmaxf(m, fabsf(p[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    // non-negative floats order like their bit patterns (fmaxf drops NaNs; an Inf input ends up as scale 2^-60, the rows
    // it touches come out non-finite and are zeroed by the guard, like the reference's GPU path)
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}
__global__ void tc_scale_kernel(const unsigned int* __restrict__ ymax, const unsigned int* __restrict__ vmax, float alpha,
                                float* __restrict__ out) {
    const float y = __uint_as_float(*ymax), w = __uint_as_float(*vmax) * fabsf(alpha);
    const float top = y * sqrtf(w);
    int e = 0;
    if (top > 0.f && isfinite(top)) {
        int ex;
        frexpf(top, &ex);          // top = f * 2^ex, f in [0.5, 1)
        e = 15 - ex;               // 2^e top in [2^14, 2^15)
    }
    e = max(-60, min(60, e));
    out[0] = ldexpf(1.f, e);
    out[1] = ldexpf(1.f, -2 * e);
}

// LOSS1: compute_loss on the item axis (the convert warps also hand sum q / sum w to the epilogue); a template parameter so
// that the common case keeps a branch-free convert loop
template <int D, bool PARTIAL, bool LOSS1>
__global__ void __launch_bounds__(THREADS, 1) als_tc_kernel(TcArgs ta) {
    static_assert(D == 128 || (D == 256 && PARTIAL), "fused row solve: d = 128; split-row mode: d = 128 or 256");
    constexpr int TILE = Cfg<D>::TILE, NACC = Cfg<D>::NACC, NF = Cfg<D>::NF, LBO = Cfg<D>::LBO;
    extern __shared__ __align__(1024) unsigned char smem_raw_[];
    Smem<D>& S = *reinterpret_cast<Smem<D>*>(smem_raw_);
    const AlsArgs& a = ta.a;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(FULL, tid >> 5, 0);

    if (tid == 0) {
        for (int i = 0; i < NR; ++i) {
            // every barrier counts WARPS, not threads: an arrive executed by 32 lanes is 32 serial barrier updates, and with
            // per-thread arrivals the hand-offs alone cost ~700 cycles per tile (measured with all the math switched off)
            mbar_init(&S.raw_full[i], N_CONV);
            mbar_init(&S.raw_empty[i], N_CONV);
        }
        for (int i = 0; i < NPG; ++i) {
            mbar_init(&S.plan_full[i], 1);
            mbar_init(&S.plan_empty[i], N_CONV);
        }
        for (int i = 0; i < NO; ++i) {
            mbar_init(&S.op_full[i], N_CONV);
            mbar_init(&S.op_empty[i], 1);
        }
        for (int i = 0; i < NACC; ++i) {
            mbar_init(&S.acc_full[i], 1);
            mbar_init(&S.acc_empty[i], 4);
            for (int b = 0; b < 4; ++b) mbar_init(&S.d_full[i][b], 1);
        }
        for (int i = 0; i < NXS; ++i) {
            mbar_init(&S.x_full[i], 4);
            S.badrow[i] = 0;
        }
        mbar_init_fence();
    }
    if (warp == W_MMA) tmem_alloc(&S.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    const float scale = __ldg(a.tc_scales), inv2 = __ldg(a.tc_scales + 1);

    // G + reg I -> tensor memory columns [0, D) (thread j of the epilogue holds matrix row j)
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

    if (warp == W_PLAN) {
        // ================= planner: gather plans =================
        uint32_t gsl = 0, rph = 0, pe = 0;   // group slot, phase of plan_empty, tile within the group
        // begin a tile: its plan slot (the group's slots are claimed when its first tile starts)
        auto tile_begin = [&]() -> uint32_t {
            if (pe == 0) mbar_wait_idle(&S.plan_empty[gsl], rph ^ 1u);
            return gsl * PT + pe;
        };
        // finish a tile (after __syncwarp): the group is published with its last tile
        auto tile_end = [&]() {
            if (++pe == PT) {
                if (lane == 0) mbar_arrive(&S.plan_full[gsl]);
                pe = 0;
                if (++gsl == NPG) { gsl = 0; rph ^= 1u; }
            }
        };
        auto emit = [&](unsigned mask, uint32_t flags, int32_t key, float w) {
            const int cnt = __popc(mask);
            const int slot = __popc(mask & ((1u << lane) - 1u));
            const uint32_t rs = tile_begin();
            if ((mask >> lane) & 1u) {
                S.keys[rs][slot] = key;
                S.sws[rs][slot] = sqrtf(fabsf(w)) * scale;
                S.wv[rs][slot] = w;
            } else {   // the lanes without an entry zero the scale / weight of the unused slots cnt .. TILE-1
                const int zslot = cnt + lane - slot;
                if (zslot < TILE) {
                    S.sws[rs][zslot] = 0.f;
                    S.wv[rs][zslot] = 0.f;
                }
            }
            if (lane == 0) S.meta_raw[rs] = (uint32_t)cnt | flags;
            __syncwarp();
            tile_end();
        };
        // Software pipeline over the CTA's items.  Every level of the dependent load chain  item -> row offsets -> entries
        // is issued whole rows ahead of its use (a gathered tile takes ~700 cycles to issue, a global load ~1-2 thousand
        // cycles under load: with a one-tile look-ahead every tile paid that latency):
        //   item i+4: row id            item i+3: offsets            item i+2: first super-chunk of entries (keys, weights)
        // and inside a long row the next super-chunk (SC tiles) is loaded while the current one is being emitted.
        // The planner is one warp on the critical path of every tile (with all the math switched off the kernel still
        // needed ~1000 cycles per tile for the plan hand-offs alone): inside a row everything is 32-bit, and a super-chunk
        // without negative weights -- the normal case -- takes the short path: slot == lane, no ballots, no prefix sums.
        constexpr int SC = 4, SCN = SC * TILE;
        const bool inl = lane < TILE;
        auto load_id = [&](int64_t it, int& row, int& chunk) {
            row = -1;
            chunk = 0;
            if (it < nitems) {
                if (PARTIAL) {
                    const int32_t* p = ta.items + 3 * (a.row_begin + it);
                    row = p[0];
                    chunk = p[1];
                } else {
                    row = a.row_list[a.row_begin + it];
                }
            }
        };
        auto load_span = [&](int row, int chunk, const int32_t*& kp, const float*& vp, int& n) {
            kp = a.keys;
            vp = a.vals;
            n = 0;
            if (row >= 0) {
                const int64_t rb = row == 0 ? 0 : a.indptr[row - 1];
                const int64_t rn = a.indptr[row] - rb;
                int64_t beg = rb;
                n = (int)rn;            // fused rows are binned below 2^31 entries, chunks are `split` long
                if (PARTIAL) {
                    beg = rb + (int64_t)chunk * ta.split;
                    n = (int)min(ta.split, rn - (int64_t)chunk * ta.split);
                }
                kp = a.keys + (beg - a.shift);
                vp = a.vals + (beg - a.shift);
            }
        };
        auto load_sc = [&](const int32_t* kp, const float* vp, int n, int c0, int32_t (&key)[SC], float (&w)[SC]) {
#pragma unroll
            for (int s = 0; s < SC; ++s) {
                const int idx = c0 + s * TILE + lane;
                const bool ok = inl && idx < n;
                key[s] = ok ? kp[idx] : 0;
                w[s] = ok ? vp[idx] * a.alpha : 0.f;
            }
        };
        auto emit_sc = [&](int n, int c0, const int32_t (&key)[SC], const float (&w)[SC]) {
            bool neg = false;
#pragma unroll
            for (int s = 0; s < SC; ++s) neg |= w[s] < 0.f;
            if (!__any_sync(FULL, neg)) {
#pragma unroll
                for (int s = 0; s < SC; ++s) {
                    const int t0 = c0 + s * TILE;
                    if (t0 < n) {
                        const uint32_t meta = (uint32_t)min(TILE, n - t0) | (t0 == 0 ? F_FIRST : 0u) | (t0 + TILE >= n ? F_LAST : 0u);
                        const uint32_t rs = tile_begin();
                        if (inl) {   // lanes beyond the row's end carry key 0, weight 0: the unused slots are zero-filled
                            float sq;
                            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(w[s]));
                            S.keys[rs][lane] = key[s];
                            S.sws[rs][lane] = sq * scale;
                            S.wv[rs][lane] = w[s];
                        }
                        if (lane == 0) S.meta_raw[rs] = meta;
                        __syncwarp();
                        tile_end();
                    }
                }
            } else {
#pragma unroll
                for (int s = 0; s < SC; ++s) {
                    const int t0 = c0 + s * TILE;
                    if (t0 < n) {
                        const bool valid = inl && t0 + lane < n;
                        const unsigned pm = __ballot_sync(FULL, valid && !(w[s] < 0.f));
                        const unsigned nm = __ballot_sync(FULL, valid && (w[s] < 0.f));
                        const bool lastc = t0 + TILE >= n;
                        uint32_t fl = (t0 == 0 ? F_FIRST : 0u);
                        if (pm) {
                            emit(pm, fl | ((lastc && !nm) ? F_LAST : 0u), key[s], w[s]);
                            fl = 0u;
                        }
                        if (nm) emit(nm, fl | F_NEG | (lastc ? F_LAST : 0u), key[s], w[s]);
                    }
                }
            }
        };
        const int32_t *kp0, *kp1, *kp2, *kp3;
        const float *vp0, *vp1, *vp2, *vp3;
        int n0, n1, n2, n3;
        int row_t, ch_t, row4, ch4;
        int32_t k0[SC], k1[SC], k2[SC], nk[SC];
        float w0[SC], w1[SC], w2[SC], nw[SC];
        load_id(my_first, row_t, ch_t);
        load_span(row_t, ch_t, kp0, vp0, n0);
        load_sc(kp0, vp0, n0, 0, k0, w0);
        load_id(my_first + stride, row_t, ch_t);
        load_span(row_t, ch_t, kp1, vp1, n1);
        load_sc(kp1, vp1, n1, 0, k1, w1);
        load_id(my_first + 2 * stride, row_t, ch_t);
        load_span(row_t, ch_t, kp2, vp2, n2);
        load_id(my_first + 3 * stride, row4, ch4);
        for (int64_t it = my_first; it < nitems; it += stride) {
            // issue the look-ahead loads (their results are first touched one item later)
            load_sc(kp2, vp2, n2, 0, k2, w2);
            load_span(row4, ch4, kp3, vp3, n3);
            load_id(it + 4 * stride, row4, ch4);
            for (int c0 = 0; c0 < n0; c0 += SCN) {
                const bool more = c0 + SCN < n0;
                if (more) load_sc(kp0, vp0, n0, c0 + SCN, nk, nw);
                emit_sc(n0, c0, k0, w0);
                if (more) {
#pragma unroll
                    for (int s = 0; s < SC; ++s) { k0[s] = nk[s]; w0[s] = nw[s]; }
                }
            }
#pragma unroll
            for (int s = 0; s < SC; ++s) { k0[s] = k1[s]; w0[s] = w1[s]; k1[s] = k2[s]; w1[s] = w2[s]; }
            kp0 = kp1; vp0 = vp1; n0 = n1; kp1 = kp2; vp1 = vp2; n1 = n2; kp2 = kp3; vp2 = vp3; n2 = n3;
        }
        // stop marker: fills the rest of the current group (or one more group)
        do {
            const uint32_t rs = tile_begin();
            if (lane == 0) S.meta_raw[rs] = F_STOP;
            __syncwarp();
            tile_end();
        } while (pe != 0);
    } else if (warp == W_MMA) {
        // ================= MMA issue =================
        uint32_t os = 0, oph = 0, acc = 0, aph = 0;
        const uint32_t idesc = idesc_f16_k(128, 128), idesc_w = idesc_f16_k(128, 256);   // d = 256: N = 256
        bool row_open = false, stop = false;
        while (!stop) {
            mbar_wait(&S.op_full[os], oph);
            fence_proxy_async_smem();   // the convert warps' generic-proxy operand stores -> visible to the tensor core's reads
            tc_fence_after();
#pragma unroll 1
            for (int e = 0; e < PT; ++e) {
                const uint32_t meta = S.meta_op[os][e];
                if (meta & F_STOP) {
                    stop = true;
                    break;
                }
                if (meta & F_FIRST) {
                    mbar_wait(&S.acc_empty[acc], aph ^ 1u);
                    tc_fence_after();
                    row_open = false;
                }
                if (lane == 0) {
                    const int ksteps = (int)(meta & 0xffu);
                    const uint32_t neg = (meta & F_NEG) ? IDESC_NEGATE_A : 0u;
                    const uint32_t hi = s32(&S.op[os][e][0][0]), lo = s32(&S.op[os][e][1][0]);
                    for (int ks = 0; ks < ((ta.debug & 2) ? 0 : ksteps); ++ks) {
                        const uint32_t acc0 = (row_open || ks > 0) ? 1u : 0u;
                        // K = 16 = two 8-k chunks: LBO apart; neighbouring 8-feature core matrices 128 B apart (SBO)
                        const uint64_t dh = smem_desc(hi + ks * 2 * LBO, LBO, 128);
                        const uint64_t dl = smem_desc(lo + ks * 2 * LBO, LBO, 128);
                        if (D == 128) {
                            const uint32_t dcol = tmem + D * (1 + acc);
                            mma_f16(dcol, dh, dh, idesc | neg, acc0);
                            mma_f16(dcol, dh, dl, idesc | neg, 1u);
                            mma_f16(dcol, dl, dh, idesc | neg, 1u);
                        } else {
                            // rows 0..127 x columns 0..255 -> tensor-memory columns [0, 256)
                            mma_f16(tmem, dh, dh, idesc_w | neg, acc0);
                            mma_f16(tmem, dh, dl, idesc_w | neg, 1u);
                            mma_f16(tmem, dl, dh, idesc_w | neg, 1u);
                            // rows 128..255 x columns 128..255 -> tensor-memory columns [256, 384): features 128.. start 16
                            // core matrices (2048 B) into the slab
                            const uint64_t eh = smem_desc(hi + ks * 2 * LBO + 2048, LBO, 128);
                            const uint64_t el = smem_desc(lo + ks * 2 * LBO + 2048, LBO, 128);
                            mma_f16(tmem + 256, eh, eh, idesc | neg, acc0);
                            mma_f16(tmem + 256, eh, el, idesc | neg, 1u);
                            mma_f16(tmem + 256, el, eh, idesc | neg, 1u);
                        }
                    }
                    if (meta & F_LAST) mma_commit(&S.acc_full[acc]);
                }
                __syncwarp();
                row_open = true;
                if (meta & F_LAST) {
                    if (++acc == NACC) { acc = 0; aph ^= 1u; }
                }
            }
            if (lane == 0) mma_commit(&S.op_empty[os]);   // (after a stop nobody waits for it any more)
            __syncwarp();
            if (++os == NO) { os = 0; oph ^= 1u; }
        }
    } else if (warp >= W_CONV && warp < W_CONV + N_CONV) {
        // ================= convert: raw fp32 -> scaled fp16 head/tail operand slabs =================
        // thread ct owns features m = ct + 128 f: a warp reads 32 consecutive floats of one gathered row (conflict-free)
        // and writes 32 consecutive 16-byte groups of a core-matrix column (conflict-free).
        const int cta = tid - W_CONV * 32;
        const int ct = cta & 127;          // feature(s) of this thread
        const int kh = cta >> 7;           // which part of a tile's entries this convert set takes
        constexpr int CHS = TILE / 8 / KH; // 8-entry chunks per tile and set
        uint32_t rs = 0, rph = 0, os = 0, oph = 0, bslot = 0;
        float2 bacc[NF], qacc[NF];
        float wacc = 0.f;
#pragma unroll
        for (int f = 0; f < NF; ++f) { bacc[f] = make_float2(0.f, 0.f); qacc[f] = make_float2(0.f, 0.f); }
        // Gathers.  The convert warps issue the gathers themselves, GATHER_AHEAD groups before they consume the group:
        // warp cw copies the TILE / 8 rows cw, cw + 8, ... of a planned tile with one coalesced 16-byte-per-lane cp.async per
        // 512 bytes (SASS LDGSTS; a whole warp instruction moves a full row segment, and the issue cost is spread over
        // eight warps that have issue slots to spare -- the 512-byte TMA bulk copies this replaces cost ~63-100 issue
        // cycles each in a dedicated warp) as one commit group per tile; before converting tile t a warp
        // waits for its own group of tile t (cp.async.wait_group) and posts one arrival on the tile's mbarrier.
        const int cw = cta >> 5;
        uint32_t gs = 0, gph = 0, grs = 0, grph = 0;   // plan group slot / raw stage of the next group to gather
        bool plan_end = false;
        auto gather = [&]() {
            if (plan_end) {
                cp_async_commit();   // keep one commit group per iteration so that wait_group counts groups to the very end
                return;
            }
            mbar_wait(&S.plan_full[gs], gph);
            mbar_wait(&S.raw_empty[grs], grph ^ 1u);   // every convert warp is done with the group that used this stage
            const uint32_t gm0 = S.meta_raw[gs * PT], gm1 = S.meta_raw[gs * PT + 1];
            if (PT == 2 && !(ta.debug & 1) && ((gm0 | gm1) & F_STOP) == 0 && (gm0 & 0xffu) == TILE && (gm1 & 0xffu) == TILE) {
                // two full tiles (the common case): straight-line copies, no per-row predicates
                int32_t kk[PT][TILE / 8];
#pragma unroll
                for (int e = 0; e < PT; ++e)
#pragma unroll
                    for (int i = 0; i < TILE / 8; ++i) kk[e][i] = S.keys[gs * PT + e][cw + 8 * i];
#pragma unroll
                for (int e = 0; e < PT; ++e)
#pragma unroll
                    for (int i = 0; i < TILE / 8; ++i) {
                        const float* src = a.Y + (int64_t)kk[e][i] * a.ld + lane * 4;
                        float* dst = &S.raw[grs][e][(cw + 8 * i) * D + lane * 4];
#pragma unroll
                        for (int c = 0; c < D / 128; ++c) cp_async16_cg(dst + c * 128, src + c * 128);
                    }
            } else
#pragma unroll
            for (int e = 0; e < PT; ++e) {
                const uint32_t gmeta = S.meta_raw[gs * PT + e];
                if (gmeta & F_STOP) {
                    plan_end = true;
                } else {
                    const int gcnt = (ta.debug & 1) ? 0 : (int)(gmeta & 0xffu);
#pragma unroll
                    for (int i = 0; i < TILE / 8; ++i) {
                        const int slot = cw + 8 * i;
                        if (slot < gcnt) {
                            const float* src = a.Y + (int64_t)S.keys[gs * PT + e][slot] * a.ld + lane * 4;
                            float* dst = &S.raw[grs][e][slot * D + lane * 4];
#pragma unroll
                            for (int c = 0; c < D / 128; ++c) cp_async16_cg(dst + c * 128, src + c * 128);
                        }
                    }
                }
            }
            cp_async_commit();
            if (++gs == NPG) { gs = 0; gph ^= 1u; }
            if (++grs == NR) { grs = 0; grph ^= 1u; }
        };
        uint32_t cs = 0;   // plan group slot of the group being converted
        bool done = false;
        if (ta.debug & 16) {   // timing experiment: consume the plans as fast as they come, nothing else
            uint32_t ph = 0;
            for (;;) {
                mbar_wait(&S.plan_full[cs], ph);
                const bool stop = (S.meta_raw[cs * PT] | S.meta_raw[cs * PT + 1]) & F_STOP;
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.plan_empty[cs]);
                if (stop) break;
                if (++cs == NPG) { cs = 0; ph ^= 1u; }
            }
            if (cta == 0) S.meta_op[0][0] = F_STOP;
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.op_full[0]);
            done = true;
        }
#pragma unroll 1
        for (int i = 0; i < GATHER_AHEAD && !done; ++i) gather();
        while (!done) {
            gather();
            cp_async_wait_group<GATHER_AHEAD>();   // this warp's rows of the group about to be converted have landed
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.raw_full[rs]);
            mbar_wait(&S.raw_full[rs], rph);        // ... and everybody else's
            // all of this thread's values of the group are read up front (independent loads, one exposed latency per group instead
            // of one per 8-entry chunk); unused slots read stale data that the guarded path zeroes
            float qv[PT][CHS * 8 * NF];
#pragma unroll
            for (int e = 0; e < PT; ++e)
#pragma unroll
                for (int c = 0; c < CHS; ++c)
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int f = 0; f < NF; ++f)
                            qv[e][(c * 8 + i) * NF + f] = S.raw[rs][e][((kh * CHS + c) * 8 + i) * D + ct + 128 * f];
            mbar_wait(&S.op_empty[os], oph ^ 1u);
            const uint32_t cm0 = S.meta_raw[cs * PT], cm1 = S.meta_raw[cs * PT + 1];
            const bool fast2 = PT == 2 && !(ta.debug & 4) && ((cm0 | cm1) & F_STOP) == 0 && (cm0 & 0xffu) == TILE &&
                               (cm1 & 0xffu) == TILE;
            if (fast2) {
                // Two full tiles: one straight-line block (the row bookkeeping is branch-free: the accumulators are reset by
                // selects on F_FIRST, the partial sums are stored after EVERY tile and the slot only advances on F_LAST), so
                // that the compiler can interleave the four chunks' load -> scale -> split -> store chains.
#pragma unroll
                for (int e = 0; e < PT; ++e) {
                    const uint32_t meta = e == 0 ? cm0 : cm1;
                    const uint32_t psl = cs * PT + e;
                    const bool first = (meta & F_FIRST) != 0;
#pragma unroll
                    for (int f = 0; f < NF; ++f) {
                        bacc[f].x = first ? 0.f : bacc[f].x; bacc[f].y = first ? 0.f : bacc[f].y;
                        qacc[f].x = first ? 0.f : qacc[f].x; qacc[f].y = first ? 0.f : qacc[f].y;
                    }
                    wacc = first ? 0.f : wacc;
                    unsigned char* hi = &S.op[os][e][0][0];
                    unsigned char* lo = &S.op[os][e][1][0];
#pragma unroll
                    for (int c = 0; c < CHS; ++c) {
                        const int kc = kh * CHS + c, k0 = kc * 8;
                        const float4 sa = *reinterpret_cast<const float4*>(&S.sws[psl][k0]);
                        const float4 sb = *reinterpret_cast<const float4*>(&S.sws[psl][k0 + 4]);
                        const float4 wa = *reinterpret_cast<const float4*>(&S.wv[psl][k0]);
                        const float4 wb = *reinterpret_cast<const float4*>(&S.wv[psl][k0 + 4]);
                        if (LOSS1) wacc += ((wa.x + wa.y) + (wa.z + wa.w)) + ((wb.x + wb.y) + (wb.z + wb.w));
#pragma unroll
                        for (int f = 0; f < NF; ++f) {
                            const int m = ct + 128 * f;
                            const float* q = &qv[e][0];
                            const float2 q01 = make_float2(q[(c * 8 + 0) * NF + f], q[(c * 8 + 1) * NF + f]);
                            const float2 q23 = make_float2(q[(c * 8 + 2) * NF + f], q[(c * 8 + 3) * NF + f]);
                            const float2 q45 = make_float2(q[(c * 8 + 4) * NF + f], q[(c * 8 + 5) * NF + f]);
                            const float2 q67 = make_float2(q[(c * 8 + 6) * NF + f], q[(c * 8 + 7) * NF + f]);
                            uint4 h4, l4;
                            split_f16x2(f2mul(q01, make_float2(sa.x, sa.y)), h4.x, l4.x);
                            split_f16x2(f2mul(q23, make_float2(sa.z, sa.w)), h4.y, l4.y);
                            split_f16x2(f2mul(q45, make_float2(sb.x, sb.y)), h4.z, l4.z);
                            split_f16x2(f2mul(q67, make_float2(sb.z, sb.w)), h4.w, l4.w);
                            const int off = kc * LBO + (m >> 3) * 128 + (m & 7) * 16;
                            *reinterpret_cast<uint4*>(hi + off) = h4;
                            *reinterpret_cast<uint4*>(lo + off) = l4;
                            bacc[f] = f2fma(make_float2(wa.x, wa.y), q01, bacc[f]);
                            bacc[f] = f2fma(make_float2(wa.z, wa.w), q23, bacc[f]);
                            bacc[f] = f2fma(make_float2(wb.x, wb.y), q45, bacc[f]);
                            bacc[f] = f2fma(make_float2(wb.z, wb.w), q67, bacc[f]);
                            if (LOSS1) {
                                qacc[f].x += (q01.x + q23.x) + (q45.x + q67.x);
                                qacc[f].y += (q01.y + q23.y) + (q45.y + q67.y);
                            }
                        }
                    }
#pragma unroll
                    for (int f = 0; f < NF; ++f) {
                        S.bvec[bslot][kh][ct + 128 * f] = bacc[f].x + bacc[f].y;
                        if (LOSS1) S.sumq[bslot][kh][ct + 128 * f] = qacc[f].x + qacc[f].y;
                    }
                    if (LOSS1 && ct == 0) S.wsum[bslot][kh] = wacc;
                    bslot = (bslot + ((meta & F_LAST) ? 1u : 0u)) & (NBV - 1);
                    if (cta == 0) S.meta_op[os][e] = (uint32_t)(TILE / 16) | (meta & (F_FIRST | F_LAST | F_NEG));
                }
            } else
#pragma unroll
            for (int e = 0; e < PT; ++e) {
                if (done) continue;
                const uint32_t psl = cs * PT + e;   // the tile's plan slot
                const uint32_t meta = S.meta_raw[psl];
                if (meta & F_STOP) {
                    if (cta == 0) S.meta_op[os][e] = F_STOP;
                    done = true;
                    continue;
                }
                const int cnt = (int)(meta & 0xffu), ksteps = (cnt + 15) >> 4;
                if (meta & F_FIRST) {
#pragma unroll
                    for (int f = 0; f < NF; ++f) { bacc[f] = make_float2(0.f, 0.f); qacc[f] = make_float2(0.f, 0.f); }
                    wacc = 0.f;
                }
                unsigned char* hi = &S.op[os][e][0][0];
                unsigned char* lo = &S.op[os][e][1][0];
                // one chunk = 8 consecutive entries k of this thread's feature(s): a 16-byte group of the head and of the
                // tail slab.  GUARD: the tile is not full -- slots >= cnt hold stale rows (scale and weight 0 from the
                // planner; the value is zeroed as well so that a stale Inf/NaN cannot leak into an unrelated row).
                auto chunk = [&](auto guard, const int c) {
                    constexpr bool GUARD = decltype(guard)::value;
                    const int kc = kh * CHS + c, k0 = kc * 8;
                    const float4 sa = *reinterpret_cast<const float4*>(&S.sws[psl][k0]);
                    const float4 sb = *reinterpret_cast<const float4*>(&S.sws[psl][k0 + 4]);
                    const float4 wa = *reinterpret_cast<const float4*>(&S.wv[psl][k0]);
                    const float4 wb = *reinterpret_cast<const float4*>(&S.wv[psl][k0 + 4]);
                    if (LOSS1) wacc += ((wa.x + wa.y) + (wa.z + wa.w)) + ((wb.x + wb.y) + (wb.z + wb.w));
#pragma unroll
                    for (int f = 0; f < NF; ++f) {
                        const int m = ct + 128 * f;
                        float q[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) q[i] = qv[e][(c * 8 + i) * NF + f];
                        if (GUARD) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) q[i] = (k0 + i < cnt) ? q[i] : 0.f;
                        }
                        const float2 q01 = make_float2(q[0], q[1]), q23 = make_float2(q[2], q[3]);
                        const float2 q45 = make_float2(q[4], q[5]), q67 = make_float2(q[6], q[7]);
                        uint4 h4, l4;
                        split_f16x2(f2mul(q01, make_float2(sa.x, sa.y)), h4.x, l4.x);
                        split_f16x2(f2mul(q23, make_float2(sa.z, sa.w)), h4.y, l4.y);
                        split_f16x2(f2mul(q45, make_float2(sb.x, sb.y)), h4.z, l4.z);
                        split_f16x2(f2mul(q67, make_float2(sb.z, sb.w)), h4.w, l4.w);
                        const int off = kc * LBO + (m >> 3) * 128 + (m & 7) * 16;
                        *reinterpret_cast<uint4*>(hi + off) = h4;
                        *reinterpret_cast<uint4*>(lo + off) = l4;
                        bacc[f] = f2fma(make_float2(wa.x, wa.y), q01, bacc[f]);
                        bacc[f] = f2fma(make_float2(wa.z, wa.w), q23, bacc[f]);
                        bacc[f] = f2fma(make_float2(wb.x, wb.y), q45, bacc[f]);
                        bacc[f] = f2fma(make_float2(wb.z, wb.w), q67, bacc[f]);
                        if (LOSS1) {
                            qacc[f].x += (q[0] + q[2]) + (q[4] + q[6]);
                            qacc[f].y += (q[1] + q[3]) + (q[5] + q[7]);
                        }
                    }
                };
                if (ta.debug & 4) {
                } else if (cnt == TILE) {   // full tile: straight-line code
#pragma unroll
                    for (int c = 0; c < CHS; ++c) chunk(std::false_type{}, c);
                } else {
#pragma unroll
                    for (int c = 0; c < CHS; ++c)
                        if (kh * CHS + c < 2 * ksteps) chunk(std::true_type{}, c);
                }
                if (meta & F_LAST) {
#pragma unroll
                    for (int f = 0; f < NF; ++f) {
                        S.bvec[bslot][kh][ct + 128 * f] = bacc[f].x + bacc[f].y;
                        if (LOSS1) S.sumq[bslot][kh][ct + 128 * f] = qacc[f].x + qacc[f].y;
                    }
                    if (LOSS1 && ct == 0) S.wsum[bslot][kh] = wacc;
                    bslot = (bslot + 1) & (NBV - 1);
                }
                if (cta == 0) S.meta_op[os][e] = (uint32_t)ksteps | (meta & (F_FIRST | F_LAST | F_NEG));
            }
            // (no proxy fence here: FENCE.VIEW.ASYNC in a thread with cp.async / global loads in flight waits for them -- the
            // next groups' gathers -- so the generic->async proxy fence is executed by the MMA-issuing warp after it has
            // acquired the group through op_full)
            __syncwarp();
            if (lane == 0) {   // one arrival per warp on each of the three hand-offs of a group
                mbar_arrive(&S.raw_empty[rs]);
                mbar_arrive(&S.plan_empty[cs]);
                mbar_arrive(&S.op_full[os]);
            }
            cs = (cs + 1) & (NPG - 1);
            if (++rs == NR) { rs = 0; rph ^= 1u; }
            if (++os == NO) { os = 0; oph ^= 1u; }
        }
    } else {
        // ================= epilogue: explicit-matrix block Gauss-Seidel / CG, systolic over rows =================
        const int q = warp & 3;                  // tensor-memory lane quarter == column block owned by this warp
        const int j = q * 32 + lane;             // matrix row
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float* pv = S.pv[q];
        double l_nume = 0.0, l_deno = 0.0;
        const float tol = a.tol;
        if constexpr (PARTIAL) {
            // add every chunk's matrix / vectors to the row's scratch block (float atomics: the chunks of one row are
            // summed in arrival order); the four warps are independent here
            int row = 0, slot = 0, nrow = 0, nslot = 0;
            int64_t beg, n = 0, nbeg, nn = 0;
            if (my_first < nitems) item_info(my_first, row, beg, n, slot);
            for (int64_t seq = 0; my_first + seq * stride < nitems; ++seq) {
                const int64_t nit = my_first + (seq + 1) * stride;
                if (nit < nitems) item_info(nit, nrow, nbeg, nn, nslot);
                const uint32_t acc = (uint32_t)(seq % NACC), aph = (uint32_t)((seq / NACC) & 1);
                const uint32_t bs = (uint32_t)(seq & (NBV - 1));
                mbar_wait_idle(&S.acc_full[acc], aph);
                tc_fence_after();
                float* sc = ta.scratch + (size_t)slot * scratch_floats<D>();
                if constexpr (D == 128) {
                    const uint32_t dbase = tmem + lane_off + D * (1 + acc);
#pragma unroll 1
                    for (int c = 0; c < D / 32; ++c) {
                        float v[32];
                        tmem_ld32(dbase + c * 32, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) atomicAdd(sc + (size_t)j * D + c * 32 + i, v[i] * inv2);
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
                            atomicAdd(sc + (size_t)j * D + c * 32 + i, v[i] * inv2);
                            if (c >= 4) atomicAdd(sc + (size_t)(c * 32 + i) * D + j, v[i] * inv2);
                        }
                    }
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        float v[32];
                        tmem_ld32(tmem + lane_off + 256 + c * 32, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) atomicAdd(sc + (size_t)(128 + j) * D + 128 + c * 32 + i, v[i] * inv2);
                    }
                }
                for (int jj = j; jj < D; jj += 128) {
                    atomicAdd(sc + (size_t)D * D + jj, S.bvec[bs][0][jj] + (KH == 2 ? S.bvec[bs][KH - 1][jj] : 0.f));
                    if (LOSS1) atomicAdd(sc + (size_t)D * D + D + jj, S.sumq[bs][0][jj] + (KH == 2 ? S.sumq[bs][KH - 1][jj] : 0.f));
                }
                if (LOSS1 && j == 0) atomicAdd(sc + (size_t)D * D + 2 * D, S.wsum[bs][0] + (KH == 2 ? S.wsum[bs][KH - 1] : 0.f));
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.acc_empty[acc]);
                row = nrow; slot = nslot; n = nn;
            }
        } else {
            // Row pipeline registers: x_j of rows seq (xj), seq + 1 (x1), and the row ids of seq + 1, seq + 2; the loads
            // for seq + 2 / seq + 3 are issued one iteration before they are touched.
            auto row_of = [&](int64_t sq) -> int {
                const int64_t it = my_first + sq * stride;
                return it < nitems ? a.row_list[a.row_begin + it] : -1;
            };
            auto len_of = [&](int row) -> float {   // only for the adaptive-reg loss term
                if (row < 0 || !a.compute_loss) return 0.f;
                return (float)(a.indptr[row] - (row == 0 ? 0 : a.indptr[row - 1]));
            };
            int row = (ta.debug & 16) ? -1 : row_of(0), row1 = row_of(1), row2 = row_of(2);
            float xj = row >= 0 ? a.X[(int64_t)row * a.ld + j] : 0.f;
            float x1 = row1 >= 0 ? a.X[(int64_t)row1 * a.ld + j] : 0.f;
            float nlen = len_of(row), nlen1 = len_of(row1);
            if (row >= 0) {   // publish row 0's x
                S.xs[0][j] = xj;
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.x_full[0]);
            }
            for (int64_t seq = 0; row >= 0; ++seq) {
                // look-ahead loads
                const int row3 = row_of(seq + 3);
                const float x2 = row2 >= 0 ? a.X[(int64_t)row2 * a.ld + j] : 0.f;
                const float nlen2 = len_of(row2);
                if (row1 >= 0) {   // publish the next row's x: by the time a warp gets there everybody has
                    S.xs[(seq + 1) & (NXS - 1)][j] = x1;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.x_full[(seq + 1) & (NXS - 1)]);
                }
                const uint32_t acc = (uint32_t)(seq % NACC), aph = (uint32_t)((seq / NACC) & 1);
                const uint32_t bs = (uint32_t)(seq & (NBV - 1));
                const uint32_t xsl = (uint32_t)(seq & (NXS - 1));
                const float* xs = S.xs[xsl];
                mbar_wait(&S.x_full[xsl], (uint32_t)((seq / NXS) & 1));
                mbar_wait_idle(&S.acc_full[acc], aph);
                tc_fence_after();
                if (ta.debug & 8) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.acc_empty[acc]);
                    row = row1; row1 = row2; row2 = row3;
                    xj = x1; x1 = x2;
                    continue;
                }
                const uint32_t dbase = tmem + lane_off + 128 * (1 + acc);
                const float bj = S.bvec[bs][0][j] + (KH == 2 ? S.bvec[bs][KH - 1][j] : 0.f);
                // ---- h = (G + reg I) x + 2^-2e A x - b  (A = the accumulator); keep the diagonal block of M in registers ----
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
                        for (int i = 0; i < 32; ++i) md[i] = fmaf(dv[i], inv2, gv[i]);
                    }
                }
                hD *= inv2;
                if (a.compute_loss) {
                    // als.cc:298-321 with the pre-update row: reg*kappa*|x|^2 (both axes); item side additionally
                    // x G x + sum_obs[(1+w)(yhat-1)^2 - yhat^2] = x G x + x D x - 2 x.(b + sum q) + (n + sum w)
                    const float kappa = a.adaptive_reg ? nlen : 1.0f;
                    double t = (double)(kappa * a.reg * xj * xj);
                    if (a.axis == 1) {
                        t += (double)xj * (double)(hG - a.reg * xj) + (double)xj * (double)hD -
                             2.0 * (double)xj * ((double)bj + (double)S.sumq[bs][0][j] + (KH == 2 ? (double)S.sumq[bs][KH - 1][j] : 0.0));
                        if (j == 0) {
                            const double ws = (double)S.wsum[bs][0] + (KH == 2 ? (double)S.wsum[bs][KH - 1] : 0.0);
                            t += (double)nlen + ws;
                            l_deno += (double)a.Y_rows + ws;
                        }
                    }
                    l_nume += t;
                }
                float h = hG + hD - bj;
                // ---- fold in the deltas of the earlier blocks as they appear: h -= M[j, B] . delta_B ----
#pragma unroll 1
                for (int B = 0; B < q; ++B) {
                    float dv[32], gv[32];
                    tmem_ld32(dbase + B * 32, dv);
                    tmem_ld32(tmem + lane_off + B * 32, gv);
                    mbar_wait(&S.d_full[acc][B], aph);
                    tmem_wait_ld();
                    float u0 = 0.f, u1 = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 d4 = *reinterpret_cast<const float4*>(&S.dl[acc][B][i]);
                        u0 = fmaf(fmaf(dv[i], inv2, gv[i]), d4.x, u0);
                        u1 = fmaf(fmaf(dv[i + 1], inv2, gv[i + 1]), d4.y, u1);
                        u0 = fmaf(fmaf(dv[i + 2], inv2, gv[i + 2]), d4.z, u0);
                        u1 = fmaf(fmaf(dv[i + 3], inv2, gv[i + 3]), d4.w, u1);
                    }
                    h -= u0 + u1;
                }
                // this warp's reads of the accumulator are done
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.acc_empty[acc]);
                // ---- 3-step CG on the own diagonal block (als.cc:324-345) ----
                float xv = 0.f;
                {
                    float r = h, p = h;
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
                }
                float v = xj - xv;                      // als.cc:346
                const bool badw = __any_sync(FULL, !isfinite(v));
                // write the own block (and the peers' replicas, fused exchange) BEFORE publishing the delta: the warp of the
                // last block may overwrite the row (NaN/Inf guard, cf. als.cu:116-120) and must come after these stores
                a.X[(int64_t)row * a.ld + j] = v;
                for (int pr = 0; pr < a.n_peer; ++pr) a.peerX[pr][(int64_t)row * a.ld + j] = v;
                if (q < 3) {
                    S.dl[acc][q][lane] = xv;
                    if (badw && lane == 0) atomicOr(&S.badrow[xsl], 1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.d_full[acc][q]);
                }
                // the warp of the last block has seen every earlier block's flag (set before that block's delta was
                // published) and zeroes the whole row if any block came out non-finite
                if (q == 3) {
                    const bool bad = badw || (S.badrow[xsl] != 0);
                    if (bad) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            a.X[(int64_t)row * a.ld + c * 32 + lane] = 0.f;
                            for (int pr = 0; pr < a.n_peer; ++pr) a.peerX[pr][(int64_t)row * a.ld + c * 32 + lane] = 0.f;
                        }
                        __syncwarp();
                        if (lane == 0) S.badrow[xsl] = 0;
                    }
                }
                row = row1; row1 = row2; row2 = row3;
                xj = x1; x1 = x2;
                nlen = nlen1; nlen1 = nlen2;
            }
            if (a.loss && a.compute_loss) {
                l_nume = warp_sum_d(l_nume);
                l_deno = warp_sum_d(l_deno);
                if (lane == 0 && (l_nume != 0.0 || l_deno != 0.0)) {
                    atomicAdd(a.loss, l_nume);
                    atomicAdd(a.loss + 1, l_deno);
                }
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

// device-side operand scale of the next tensor-core launches: maxes[0] = bits of max|Y| (noted at Gram time),
// maxes[1] = bits of max|v| over the values of this launch
inline int tc_note_absmax(const float* p, size_t n, unsigned int* out, int num_sms, cudaStream_t st) {
    BFL_CUDA(cudaMemsetAsync(out, 0, sizeof(unsigned int), st));
    if (n == 0) return BFL_OK;
    const int grid = (int)std::min<size_t>((n / 4 + 255) / 256 + 1, (size_t)num_sms * 8);
    tc_absmax_kernel<<<grid, 256, 0, st>>>(p, n, out);
    BFL_LAUNCHED();
    return BFL_OK;
}
inline int tc_update_scale(const unsigned int* ymax, const unsigned int* vmax, float alpha, float* scales, cudaStream_t st) {
    tc_scale_kernel<<<1, 1, 0, st>>>(ymax, vmax, alpha, scales);
    BFL_LAUNCHED();
    return BFL_OK;
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
    if (!a.tc_scales) BFL_FAIL(BFL_ERR_STATE, "tensor-core ALS kernel: operand scale not prepared");
    TcArgs ta;
    ta.a = a;
    ta.a.row_begin = 0;
    ta.a.row_end = nitems;
    ta.items = items;
    ta.scratch = scratch;
    ta.split = split;
    ta.debug = 0;
    BFL_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * scratch_floats<D>() * (size_t)nslots, st));
    const size_t smem = sizeof(Smem<D>);
    const int grid = (int)std::min<int64_t>(nitems, (int64_t)num_sms);
    if (a.compute_loss && a.axis == 1) {
        BFL_CUDA(cudaFuncSetAttribute(als_tc_kernel<D, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        als_tc_kernel<D, true, true><<<grid, THREADS, smem, st>>>(ta);
    } else {
        BFL_CUDA(cudaFuncSetAttribute(als_tc_kernel<D, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        als_tc_kernel<D, true, false><<<grid, THREADS, smem, st>>>(ta);
    }
    BFL_LAUNCHED();
    return BFL_OK;
}

// solves the rows a.row_list[a.row_begin .. a.row_end) (any length > 0) with the fused tensor-core kernel
inline int tc_launch(const AlsArgs& a, int num_sms, cudaStream_t st) {
    const int64_t nrows = a.row_end - a.row_begin;
    if (nrows <= 0) return BFL_OK;
    if (!a.tc_scales) BFL_FAIL(BFL_ERR_STATE, "tensor-core ALS kernel: operand scale not prepared");
    TcArgs ta;
    ta.a = a;
    ta.items = nullptr;
    ta.scratch = nullptr;
    ta.split = 0;
    ta.debug = getenv("BFL_TC_DEBUG") ? atoi(getenv("BFL_TC_DEBUG")) : 0;
    const size_t smem = sizeof(Smem<128>);
    const int grid = (int)std::min<int64_t>(nrows, (int64_t)num_sms);
    if (a.compute_loss && a.axis == 1) {
        BFL_CUDA(cudaFuncSetAttribute(als_tc_kernel<128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        als_tc_kernel<128, false, true><<<grid, THREADS, smem, st>>>(ta);
    } else {
        BFL_CUDA(cudaFuncSetAttribute(als_tc_kernel<128, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        als_tc_kernel<128, false, false><<<grid, THREADS, smem, st>>>(ta);
    }
    BFL_LAUNCHED();
    return BFL_OK;
}

}  // namespace tc
}  // namespace bfl
