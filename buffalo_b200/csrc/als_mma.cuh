// iALS++ row-solve kernel with the per-block Gram matrix formed on the tensor cores (sm_100a).
//
// The reference runs three CG steps per 32-column block on the IMPLICIT operator
//     A = G[blk,blk] + reg I + sum_s w_s q_s q_s^T            (lib/algo_impl/als/als.cc:268-345)
// i.e. six passes over the row's gathered segments per block.  A does not depend on the row being solved, so
// this kernel forms it EXPLICITLY once per block,  sum_s (w_s q_s) q_s^T,  with mma.sync.m16n8k8 TF32 in the
// error-compensated 3xTF32 form (hi*hi + lo*hi + hi*lo, fp32 accumulate; each operand is split as u = hi + lo
// with hi = u truncated to 10 mantissa bits), and then runs the same three CG
// steps on the 32x32 matrix held in registers.  The gradient b, the Yui bookkeeping and the CG recurrences stay
// in fp32 exactly as in the SIMT kernel (als_fast.cuh), so the result differs from it only by summation order.
//
// Lane layout: lane = 4*g + t.  Of a tile of 32 gathered rows a lane owns the 8 slots c*8 + 2t + h (c = 0..3
// k-chunks, h = 0..1) and of the block's 32 columns the 4 columns 4g..4g+3 (one 16-byte load per slot).  With
// MMA row/column index i <-> block column 4*(i%8) + i/8 and k index t + 4h <-> slot 2t + h, those 8 floats are
// exactly what the lane's A fragments (both 16-row tiles, scaled by w) AND its B fragments (all four 8-column
// tiles) are made of: no shuffles.
// The accumulator fragments then hold the 4 x 8 patch  rows 4g..4g+3, columns 8t..8t+7  of A in natural order.
#pragma once
#include "als_fast.cuh"

namespace bfl {

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// sum over the 4 t-lanes (lane bits 0,1) of four values; lane t keeps value t
__device__ __forceinline__ float reduce_t4(float v0, float v1, float v2, float v3, int t) {
    const bool h2 = t & 2, h1 = t & 1;
    float k0 = h2 ? v2 : v0, k1 = h2 ? v3 : v1;
    const float s0 = h2 ? v0 : v2, s1 = h2 ? v1 : v3;
    k0 += __shfl_xor_sync(FULL, s0, 2);
    k1 += __shfl_xor_sync(FULL, s1, 2);
    return (h1 ? k1 : k0) + __shfl_xor_sync(FULL, h1 ? k0 : k1, 1);
}

// sum over the 8 g-lanes (lane bits 2..4) of eight values; lane g keeps value g
__device__ __forceinline__ float reduce_g8(const float (&d)[8], int g) {
    const bool h4 = g & 4, h2 = g & 2, h1 = g & 1;
    float k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) k[i] = (h4 ? d[4 + i] : d[i]) + __shfl_xor_sync(FULL, h4 ? d[i] : d[4 + i], 16);
    float m[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) m[i] = (h2 ? k[2 + i] : k[i]) + __shfl_xor_sync(FULL, h2 ? k[i] : k[2 + i], 8);
    return (h1 ? m[1] : m[0]) + __shfl_xor_sync(FULL, h1 ? m[0] : m[1], 4);
}

__device__ __forceinline__ float2 lds2(const float* p) { return *reinterpret_cast<const float2*>(p); }
// one k-chunk of a lane from its two staging cells (slot x at src, slot y 512 B further): x01 y01 x23 y23
__device__ __forceinline__ void load_chunk(float2 (&qc)[4], const float* src) {
    qc[0] = lds2(src);
    qc[1] = lds2(src + 128);
    qc[2] = lds2(src + 2);
    qc[3] = lds2(src + 130);
}

// dynamic smem layout:
//   [GSM ? Gs[D*(D+4)] : -] | staging: 16 warps x (K+KS) tiles x 8 cells x 32 lanes x 16 B |
//   per team: xs[D] dl[32] pw[32] bp[W*32] yui[cap] wv[cap] ks[cap]
// The first 4 KB of a warp's staging area doubles as its partial-A buffer between the MMA phase and the team
// reduction; the second 4 KB of the team's first warp holds the reduced A (W >= 4).
__host__ __device__ inline size_t mma_team_floats(int D, int W, int cap) { return (size_t)D + 64 + 32 * W + 3 * (size_t)cap; }
__host__ __device__ inline size_t mma_smem_bytes(int D, int W, int K, int KS, bool gsm, int cap) {
    return sizeof(float) * ((gsm ? (size_t)D * (D + 4) : 0) + (size_t)FAST_WARPS * (K + KS) * 1024 +
                            (FAST_WARPS / W) * mma_team_floats(D, W, cap));
}

template <int W, int K, int KS, bool GSM>
__global__ void __launch_bounds__(FAST_THREADS, 1) als_ialspp_mma_kernel(AlsArgs a, int cap) {
    static_assert(W == 1 || K >= 2, "the partial-A buffer needs two staging tiles per warp");
    extern __shared__ __align__(16) float smem[];
    constexpr int TEAMS = FAST_WARPS / W;
    constexpr int KT = K + KS;
    const int D = a.D, ld = a.ld, GP = GSM ? D + 4 : D, NB = D >> 5;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(FULL, tid >> 5, 0);   // warp-uniform for the compiler (see als_fast.cuh)
    const int team = warp / W, wt = warp % W;
    const int g = lane >> 2, t = lane & 3;
    float* Gs = smem;
    float* stg_all = smem + (GSM ? (size_t)D * (D + 4) : 0);
    float* stg = stg_all + (size_t)warp * KT * 1024;
    float* tb = stg_all + (size_t)FAST_WARPS * KT * 1024 + (size_t)team * mma_team_floats(D, W, cap);
    float* xs = tb;
    float* dl = xs + D;       // the block's solution delta, published by the solver warp
    float* pw = dl + 32;      // solver warp's CG direction
    float* bp = pw + 32;      // [W][32] per-warp partial gradients
    float* yui = bp + 32 * W;
    float* wv = yui + cap;
    int32_t* ks = reinterpret_cast<int32_t*>(wv + cap);
    const float* Gp = GSM ? Gs : a.G;
    float* team_stg = stg_all + (size_t)(team * W) * KT * 1024;   // staging area of the team's first warp
    float* afin = team_stg + 1024;

    if (GSM) {
        for (int e = tid * 4; e < D * D; e += FAST_THREADS * 4) {
            const float4 gv = ldg4(a.G + e);
            const int r = e / D, c = e - r * D;
            *reinterpret_cast<float4*>(Gs + r * GP + c) = gv;
        }
        __syncthreads();
    }

    double l_nume = 0.0, l_deno = 0.0;
    const float tol = a.tol;
    const int64_t stride = (int64_t)gridDim.x * TEAMS;

    // async copies of this lane's cells of column block B; reg_tiles / smem_tiles select which tiles
    auto stage_tiles = [&](int B, int ntiles, bool reg_tiles, bool smem_tiles) {
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
            if (kk < K ? !reg_tiles : !smem_tiles) continue;
            const int T = wt + kk * W;
            if (kk < K || T < ntiles) {   // register tiles: always (padded slots gather a valid row, weight 0)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int2 kp = *reinterpret_cast<const int2*>(ks + T * 32 + c * 8 + 2 * t);
                    float* dst = stg + ((kk * 8 + c * 2) * 32 + lane) * 4;
                    cp_async16(dst, a.Y + (int64_t)kp.x * ld + B * 32 + 4 * g);
                    cp_async16(dst + 128, a.Y + (int64_t)kp.y * ld + B * 32 + 4 * g);
                }
            }
        }
        cp_async_commit();
    };

    for (int64_t ri = a.row_begin + (int64_t)blockIdx.x * TEAMS + team; ri < a.row_end; ri += stride) {
        const int row = __shfl_sync(FULL, a.row_list[ri], 0);
        const int64_t beg = row == 0 ? 0 : a.indptr[row - 1];
        const int n = __shfl_sync(FULL, (int)(a.indptr[row] - beg), 0);
        const int ntiles = (n + 31) >> 5;
        float* xrow = a.X + (int64_t)row * ld;
        team_sync<W>(team);  // the previous row's readers are done with the team's smem
        {
            const int32_t k0 = a.keys[beg - a.shift];
            for (int c = wt * 32 + lane; c < cap; c += 32 * W) {
                const bool ok = c < n;
                const float w = ok ? a.vals[beg - a.shift + c] * a.alpha : 0.f;
                ks[c] = ok ? a.keys[beg - a.shift + c] : k0;   // padded slots gather a valid row ...
                wv[c] = w;                                       // ... with weight 0
                if (!ok) yui[c] = 0.f;
            }
        }
        for (int j = wt * 32 + lane; j < D; j += 32 * W) xs[j] = xrow[j];
        team_sync<W>(team);
        stage_tiles(0, ntiles, true, true);   // block 0's segments fly while the Yui pass streams the rows

        // ---- Yui = x . q_c over all D columns (als.cc:256-266), loss pieces with the pre-update row ----
        for (int T = wt; T < ntiles; T += W) {
            const float* rowp[8];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int2 kp = *reinterpret_cast<const int2*>(ks + T * 32 + c * 8 + 2 * t);
                rowp[2 * c] = a.Y + (int64_t)kp.x * ld + 4 * g;
                rowp[2 * c + 1] = a.Y + (int64_t)kp.y * ld + 4 * g;
            }
            float2 part[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) part[v] = make_float2(0.f, 0.f);
            for (int B = 0; B < NB; ++B) {
                float4 qv[8];
#pragma unroll
                for (int v = 0; v < 8; ++v) qv[v] = ldg4(rowp[v] + B * 32);
                const float4 x4 = lds4(xs + B * 32 + 4 * g);
                const float2 x01 = make_float2(x4.x, x4.y), x23 = make_float2(x4.z, x4.w);
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    part[v] = ffma2(make_float2(qv[v].x, qv[v].y), x01, part[v]);
                    part[v] = ffma2(make_float2(qv[v].z, qv[v].w), x23, part[v]);
                }
            }
            float d[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) d[v] = part[v].x + part[v].y;
            const float dot = reduce_g8(d, g);   // lane g now holds the dot of its slot number g
            const int slot = T * 32 + (g >> 1) * 8 + 2 * t + (g & 1);
            if (slot < n) {
                yui[slot] = dot;
                if (a.compute_loss && a.axis == 1) {  // als.cc:310-315
                    const float av = wv[slot];
                    l_nume += -(double)(dot * dot) + (double)((dot - 1.f) * (dot - 1.f)) * (1.0 + (double)av);
                    l_deno += (double)av;
                }
            }
        }
        // pull the next row's gathered rows towards L2 while this row is being solved
        if (W <= 8 && ri + stride < a.row_end) {   // long rows: the in-flight working set already fills L2
            const int64_t row2 = a.row_list[ri + stride];
            const int64_t beg2 = row2 == 0 ? 0 : a.indptr[row2 - 1];
            const int n2 = (int)(a.indptr[row2] - beg2);
            for (int c = wt * 32 + lane; c < n2 * NB; c += 32 * W) {
                const int s2 = c / NB, l2 = c - s2 * NB;
                prefetch_l2(a.Y + (int64_t)a.keys[beg2 - a.shift + s2] * ld + l2 * 32);
            }
        }
        if (a.compute_loss) {
            // reg * kappa * |x|^2 (als.cc:319-321) and, item side, x G x (als.cc:298-301); team-strided over j
            float xx = 0.f, xgx = 0.f;
            for (int j = wt * 32 + lane; j < D; j += 32 * W) {
                const float xj = xs[j];
                xx += xj * xj;
                if (a.axis == 1) {
                    float s = 0.f;
                    for (int k = 0; k < D; ++k) s = fmaf(xs[k], Gp[k * GP + j], s);
                    xgx += xj * s;
                }
            }
            xx = warp_sum(xx);
            xgx = warp_sum(xgx);
            if (lane == 0) {
                l_nume += (double)((a.adaptive_reg ? (float)n : 1.0f) * a.reg * xx);
                if (a.axis == 1) {
                    l_nume += (double)xgx;
                    if (wt == 0) l_deno += (double)a.Y_rows;
                }
            }
        }
        team_sync<W>(team);

        // ---- column blocks (als.cc:268-352) ----
        for (int B = 0; B < NB; ++B) {
            // per k-chunk four register pairs in the order x01 y01 x23 y23 (x, y = the lane's two slots): the A
            // fragments of both 16-row tiles are then two aligned register quads as loaded
            float2 q[K][4][4];
            cp_async_wait_all();   // each lane reads back only what it copied itself: no barrier needed
#pragma unroll
            for (int kk = 0; kk < K; ++kk)
#pragma unroll
                for (int c = 0; c < 4; ++c) load_chunk(q[kk][c], stg + ((kk * 8 + c * 2) * 32 + lane) * 4);
            if (W == 1 && B + 1 < NB) stage_tiles(B + 1, ntiles, true, false);   // prefetch behind the math

            // A accumulators (32 x 32 over the warp): the solver warp starts from G[blk,blk]
            float acc[2][4][4];
            if (wt == 0) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int er = 0; er < 2; ++er)
#pragma unroll
                        for (int ec = 0; ec < 2; ++ec) {
                            const float* gr = Gp + (size_t)(B * 32 + 4 * g + 2 * mt + er) * GP + B * 32 + 8 * t + 4 * ec;
                            const float4 gv = GSM ? lds4(gr) : ldg4(gr);
                            acc[mt][0][2 * er + ec] = gv.x;
                            acc[mt][1][2 * er + ec] = gv.y;
                            acc[mt][2][2 * er + ec] = gv.z;
                            acc[mt][3][2 * er + ec] = gv.w;
                        }
            } else {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
            }
            float2 b01 = make_float2(0.f, 0.f), b23 = make_float2(0.f, 0.f);   // gradient partial, columns 4g..4g+3

            // one tile: b += sum (Yui - 1) w q  (als.cc:303-308) and A += sum (w q) q^T
            auto tile_ab = [&](int T, const float2(&qt)[4][4]) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int s0 = T * 32 + c * 8 + 2 * t;
                    const float2 y2 = lds2(yui + s0), w2 = lds2(wv + s0);
                    const float c0 = fmaf(y2.x, w2.x, -w2.x), c1 = fmaf(y2.y, w2.y, -w2.y);
                    b01 = ffma2(make_float2(c0, c0), qt[c][0], b01);
                    b23 = ffma2(make_float2(c0, c0), qt[c][2], b23);
                    b01 = ffma2(make_float2(c1, c1), qt[c][1], b01);
                    b23 = ffma2(make_float2(c1, c1), qt[c][3], b23);
                    // A operand: q (fragment quads = the loaded pairs), B operand: w q (computed into pairs); each is
                    // split as hi + lo with hi = the value truncated to TF32 (the tensor core ignores the low 13 bits)
                    const float qa[2][4] = {{qt[c][0].x, qt[c][0].y, qt[c][2].x, qt[c][2].y},
                                            {qt[c][1].x, qt[c][1].y, qt[c][3].x, qt[c][3].y}};
                    const float wh[2] = {w2.x, w2.y};
                    uint32_t ah[2][4], al[2][4], bh[2][4], bl[2][4];
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            ah[h][m] = __float_as_uint(qa[h][m]) & 0xffffe000u;
                            al[h][m] = __float_as_uint(qa[h][m] - __uint_as_float(ah[h][m]));
                            const float u = wh[h] * qa[h][m];
                            bh[h][m] = __float_as_uint(u) & 0xffffe000u;
                            bl[h][m] = __float_as_uint(u - __uint_as_float(bh[h][m]));
                        }
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            mma_tf32(acc[mt][nt], al[0][2 * mt], al[0][2 * mt + 1], al[1][2 * mt], al[1][2 * mt + 1],
                                     bh[0][nt], bh[1][nt]);
                            mma_tf32(acc[mt][nt], ah[0][2 * mt], ah[0][2 * mt + 1], ah[1][2 * mt], ah[1][2 * mt + 1],
                                     bl[0][nt], bl[1][nt]);
                            mma_tf32(acc[mt][nt], ah[0][2 * mt], ah[0][2 * mt + 1], ah[1][2 * mt], ah[1][2 * mt + 1],
                                     bh[0][nt], bh[1][nt]);
                        }
                }
            };
#pragma unroll
            for (int kk = 0; kk < K; ++kk) tile_ab(wt + kk * W, q[kk]);
#pragma unroll
            for (int kk = K; kk < KT; ++kk) {
                const int T = wt + kk * W;
                if (T < ntiles) {
                    float2 qs[4][4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) load_chunk(qs[c], stg + ((kk * 8 + c * 2) * 32 + lane) * 4);
                    tile_ab(T, qs);
                }
            }
            // dense part of the gradient, x G[:, blk] (als.cc:296): rows i = 4j + t, team-strided over j
            for (int j = wt; j < (D >> 2); j += W) {
                const int i = 4 * j + t;
                const float xi = xs[i];
                const float* gr = Gp + (size_t)i * GP + B * 32 + 4 * g;
                const float4 gv = GSM ? lds4(gr) : ldg4(gr);
                b01 = ffma2(make_float2(xi, xi), make_float2(gv.x, gv.y), b01);
                b23 = ffma2(make_float2(xi, xi), make_float2(gv.z, gv.w), b23);
            }
            float bsum = reduce_t4(b01.x, b01.y, b23.x, b23.y, t);   // lane owns column `lane` of the block

            // ---- team reduction of A and b into the solver warp (wt == 0) ----
            if (W == 2) {
                if (wt == 1) {
                    __syncwarp();
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt)
                            *reinterpret_cast<float4*>(stg + ((mt * 4 + nt) * 32 + lane) * 4) =
                                make_float4(acc[mt][nt][0], acc[mt][nt][1], acc[mt][nt][2], acc[mt][nt][3]);
                    bp[32 + lane] = bsum;
                }
                team_sync<W>(team);
                if (wt == 0) {
                    const float* ps = stg + KT * 1024;   // the partner's staging area
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            const float4 pv = lds4(ps + ((mt * 4 + nt) * 32 + lane) * 4);
                            acc[mt][nt][0] += pv.x; acc[mt][nt][1] += pv.y; acc[mt][nt][2] += pv.z; acc[mt][nt][3] += pv.w;
                        }
                    bsum += bp[32 + lane];
                }
                team_sync<W>(team);
            } else if (W >= 4) {
                __syncwarp();
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
                        *reinterpret_cast<float4*>(stg + ((mt * 4 + nt) * 32 + lane) * 4) =
                            make_float4(acc[mt][nt][0], acc[mt][nt][1], acc[mt][nt][2], acc[mt][nt][3]);
                bp[wt * 32 + lane] = bsum;
                team_sync<W>(team);
                for (int cell = wt * 32 + lane; cell < 256; cell += 32 * W) {   // every warp sums a slice of the cells
                    float4 s = lds4(team_stg + cell * 4);
#pragma unroll
                    for (int w = 1; w < W; ++w) {
                        const float4 pv = lds4(team_stg + (size_t)w * KT * 1024 + cell * 4);
                        s.x += pv.x; s.y += pv.y; s.z += pv.z; s.w += pv.w;
                    }
                    *reinterpret_cast<float4*>(afin + cell * 4) = s;
                }
                team_sync<W>(team);
                if (wt == 0) {
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            const float4 pv = lds4(afin + ((mt * 4 + nt) * 32 + lane) * 4);
                            acc[mt][nt][0] = pv.x; acc[mt][nt][1] = pv.y; acc[mt][nt][2] = pv.z; acc[mt][nt][3] = pv.w;
                        }
                    bsum = 0.f;
#pragma unroll
                    for (int w = 0; w < W; ++w) bsum += bp[w * 32 + lane];
                }
            }
            // the staging area is free again: fetch the next block's register tiles behind the solve
            if (W >= 2 && B + 1 < NB) stage_tiles(B + 1, ntiles, true, false);

            // ---- 3 CG steps on A delta = g, A = G[blk,blk] + reg I + sum w q q^T (als.cc:278,324-345) ----
            // The reference skips the solve when rsold <= tol and leaves the loop when rsnew < tol (als.cc:329,341);
            // here the three steps always run and those conditions only mask the updates (same results).
            if (wt == 0) {
                const float gv = bsum + a.reg * xs[B * 32 + lane];
                float xv = 0.f, r = gv, p = gv;
                float rsold = warp_sum(r * r);
                bool act = rsold > tol;
#pragma unroll 1
                for (int step = 0; step < 3; ++step) {
                    __syncwarp();
                    pw[lane] = p;
                    __syncwarp();
                    const float4 pL = lds4(pw + 8 * t), pH = lds4(pw + 8 * t + 4);
                    float y[2][2];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int er = 0; er < 2; ++er) {
                            float2 s = fmul2(make_float2(acc[mt][0][2 * er], acc[mt][0][2 * er + 1]), make_float2(pL.x, pH.x));
                            s = ffma2(make_float2(acc[mt][1][2 * er], acc[mt][1][2 * er + 1]), make_float2(pL.y, pH.y), s);
                            s = ffma2(make_float2(acc[mt][2][2 * er], acc[mt][2][2 * er + 1]), make_float2(pL.z, pH.z), s);
                            s = ffma2(make_float2(acc[mt][3][2 * er], acc[mt][3][2 * er + 1]), make_float2(pL.w, pH.w), s);
                            y[mt][er] = s.x + s.y;
                        }
                    const float Ap = reduce_t4(y[0][0], y[0][1], y[1][0], y[1][1], t) + a.reg * p;
                    const float pAp = warp_sum(p * Ap);
                    // als.cc:337 (no eps): fp32 division of the two floats, see als_fast.cuh
                    const float step_size = act ? __fdiv_rn(rsold, pAp) : 0.f;
                    xv = fmaf(step_size, p, xv);
                    r = fmaf(-step_size, Ap, r);
                    const float rsnew = warp_sum(r * r);
                    act = act && !(rsnew < tol);                       // als.cc:341
                    if (act) p = fmaf(__fdiv_rn(rsnew, rsold), p, r);
                    rsold = act ? rsnew : rsold;
                }
                dl[lane] = xv;
                xs[B * 32 + lane] -= xv;   // x_blk -= delta (als.cc:346)
            }
            team_sync<W>(team);

            // ---- Yui -= q_blk . delta  (als.cc:347-350); not read again after the last block ----
            if (B + 1 < NB) {
                const float2 d01 = lds2(dl + 4 * g), d23 = lds2(dl + 4 * g + 2);
                auto tile_yui = [&](int T, const float2(&qt)[4][4]) {
                    float d[8];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float2 sx = ffma2(qt[c][2], d23, fmul2(qt[c][0], d01));
                        const float2 sy = ffma2(qt[c][3], d23, fmul2(qt[c][1], d01));
                        d[2 * c] = sx.x + sx.y;
                        d[2 * c + 1] = sy.x + sy.y;
                    }
                    const float dot = reduce_g8(d, g);
                    const int slot = T * 32 + (g >> 1) * 8 + 2 * t + (g & 1);
                    if (slot < n) yui[slot] -= dot;
                };
#pragma unroll
                for (int kk = 0; kk < K; ++kk) tile_yui(wt + kk * W, q[kk]);
#pragma unroll
                for (int kk = K; kk < KT; ++kk) {
                    const int T = wt + kk * W;
                    if (T < ntiles) {
                        float2 qs[4][4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) load_chunk(qs[c], stg + ((kk * 8 + c * 2) * 32 + lane) * 4);
                        tile_yui(T, qs);
                    }
                }
                if (KS > 0) stage_tiles(B + 1, ntiles, false, true);   // smem-resident tiles are free only now
                __syncwarp();   // the next block reads Yui slots written by other lanes of this warp
            }
        }
        // NaN/Inf guard (cf. als.cu:116-120), then write the row back
        bool bad = false;
        for (int j = lane; j < D; j += 32) bad |= !isfinite(xs[j]);
        bad = __any_sync(FULL, bad);
        for (int j = wt * 32 + lane; j < D; j += 32 * W) {
            const float v = bad ? 0.f : xs[j];
            xrow[j] = v;
            // fused exchange: the same 128-byte segments go straight into the peers' replicas over NVLink
            for (int pr = 0; pr < a.n_peer; ++pr) a.peerX[pr][(int64_t)row * ld + j] = v;
        }
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

template <int W, int K, int KS, bool GSM>
int mma_launch_class(const AlsArgs& a, int cap, int num_sms, cudaStream_t st) {
    const size_t smem = mma_smem_bytes(a.D, W, K, KS, GSM, cap);
    constexpr int SMEM_MAX = 227 * 1024;
    static bool configured = false;
    if (!configured) {
        BFL_CUDA(cudaFuncSetAttribute(als_ialspp_mma_kernel<W, K, KS, GSM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SMEM_MAX));
        configured = true;
    }
    if (smem > (size_t)SMEM_MAX) BFL_FAIL(BFL_ERR_STATE, "tensor-core ALS kernel: shared memory budget exceeded");
    const int64_t nrows = a.row_end - a.row_begin;
    constexpr int TEAMS = FAST_WARPS / W;
    const int grid = (int)std::min<int64_t>((nrows + TEAMS - 1) / TEAMS, (int64_t)num_sms);
    als_ialspp_mma_kernel<W, K, KS, GSM><<<grid, FAST_THREADS, smem, st>>>(a, cap);
    BFL_LAUNCHED();
    return BFL_OK;
}

// classes 0..5 (rows up to 1536 nnz) on the tensor-core kernel; returns false for the classes it does not cover
inline bool mma_class_covered(int c) { return c >= 0 && c <= 5; }
inline int mma_launch(int c, const AlsArgs& a, int cap, int num_sms, cudaStream_t st) {
    switch (c) {
        case 0: return mma_launch_class<1, 1, 0, true>(a, cap, num_sms, st);
        case 1: return mma_launch_class<1, 2, 0, true>(a, cap, num_sms, st);
        case 2: return mma_launch_class<2, 2, 0, true>(a, cap, num_sms, st);
        case 3: return mma_launch_class<4, 2, 0, true>(a, cap, num_sms, st);
        case 4: return mma_launch_class<8, 2, 0, true>(a, cap, num_sms, st);
        case 5: return mma_launch_class<16, 2, 1, false>(a, cap, num_sms, st);
    }
    return BFL_ERR_STATE;
}

}  // namespace bfl
