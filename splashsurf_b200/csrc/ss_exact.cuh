// ss_exact.cuh -- exact pass of level-set variant 2: the ordered cubic-spline fold for the boxes certification could not
// settle (surface band, outside points near the fluid, fix-up points), one WARP per flagged brick, no CTA barrier.
//
// The values must equal the reference's bit for bit (dense_subdomains.rs:1044-1128): per grid point an ordered fold over the
// subdomain's particles in ascending global index with fixed roundings.  Per brick the warp
//   1. stages the candidate records with bulk asynchronous copies (TMA engine) like the certification kernel, plus the
//      particle index (sort key) and the AVX-remainder split of every candidate with ordinary coalesced loads;
//   2. sorts the (index, slot) keys once -- a bitonic network in its private shared-memory slice, warp-synchronous;
//   3. for every 4x4x4-point sub-box with a flagged 2x4x4 box: compacts, in sorted order, the candidates within the kernel
//      support of the sub-box, then folds them into TWO grid points per lane (k, k + 1) with packed FP32.  Each element of a
//      packed instruction is rounded exactly like its scalar form, and the operation sequence per element is the reference's:
//          d^2 = fma(dz, dz, fma(dx, dx, dy * dy));  r = sqrt(d^2);  W = the AVX kernel's fnmadd / fmadd chain;
//          phi = fma(W, V, phi)      (phi + W * V in the 8-lane remainder of the particle's stencil)
//      A candidate outside the support of one of the two points contributes an exact zero there (W := 0).
// Scalar-arithmetic tiles (sparse subdomains, simd off) and the global path use the one-point formulas twice.
// Bricks this kernel cannot take (more candidates than its slice holds) are appended to a fallback list for k_levelset.
#pragma once

#define SS_XW_WARPS 3                  // bricks in flight per CTA (3 x 13.5 KB of static shared memory)
#define SS_XW_THREADS (SS_XW_WARPS * 32)
#define SS_XW_CAP 512                  // candidates staged per brick
#define SS_XW_LIST 224                 // candidates within the support of one 4x4x4 sub-box

struct SsXwArgs {
    const uint32_t *bin_start, *bin_end;
    const float4 *rec;
    const int *ksplit;
    const uint32_t *pidx;
    const SsTile *tile_tab;
    const int2 *brick_rng;
    const uint32_t *bricks;            // flagged bricks (linear index)
    uint32_t n_bricks;
    const uint8_t *wflag;              // [batch][nb^3][16] boxes that need exact values
    float *tiles;
    uint32_t *fallback;                // bricks left to k_levelset (SS_LS_FIX): [0] = count, [1..] = linear brick indices
    unsigned long long *pairs;         // work counter (in-support evaluations), only with COUNT
};

template <int CAP, int LIST, int KEYS>
struct __align__(16) SsXwSliceT {
    float4 rec[CAP];
    unsigned long long key[KEYS];      // (particle index << 32) | slot, sorted ascending
    uint16_t ks[CAP];                  // AVX-remainder split per slot
    uint16_t list[LIST];               // slots within the support of the current sub-box, ascending particle index
    unsigned long long mbar;
    unsigned long long pad_;
};
typedef SsXwSliceT<SS_XW_CAP, SS_XW_LIST, 512> SsXwSlice;
// Dense clusters (overlapping droplets, cfg-5): the same kernel with one warp per CTA and a 108 KB slice of dynamic shared memory
#define SS_XW_BIG_CAP 4096
#define SS_XW_BIG_LIST 2048
typedef SsXwSliceT<SS_XW_BIG_CAP, SS_XW_BIG_LIST, SS_XW_BIG_CAP> SsXwSliceBig;
// in between: up to 1024 candidates (two or three overlapping bodies of fluid), one warp per CTA, 27 KB of static shared memory ->
// 8 bricks in flight per SM instead of the big variant's 2
#define SS_XW_MID_CAP 1024
#define SS_XW_MID_LIST 640
typedef SsXwSliceT<SS_XW_MID_CAP, SS_XW_MID_LIST, SS_XW_MID_CAP> SsXwSliceMid;
#ifdef SS_HOST_EMUL
static thread_local __align__(16) unsigned char ss_dyn_smem[sizeof(SsXwSliceBig)];
#else
extern __shared__ __align__(16) unsigned char ss_dyn_smem[];
#endif

// warp-synchronous bitonic sort of n (power of two) keys in shared memory
__device__ __forceinline__ void ss_warp_bitonic(unsigned long long *keys, int n, int lane) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (n >> 1); t += 32) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int hi = lo | j;
                const unsigned long long a = keys[lo], b = keys[hi];
                const bool up = ((lo & k) == 0);
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
            __syncwarp();
        }
    }
}

// packed AVX-path kernel (kernel.rs:343-378) of two radii; every element rounds like ss_kernel_avx
__device__ __forceinline__ ss_f2 ss_kernel_avx2(const SsDev &P, float ra, float rb) {
    const ss_f2 q = ss_mul2(ss_pack(ra, rb), ss_pack(P.a_hinv, P.a_hinv));
    float qa, qb;
    ss_unpack(q, qa, qb);
    const ss_f2 one_m = ss_fma2(q, ss_pack(-1.0f, -1.0f), ss_pack(1.0f, 1.0f));          // 1 - q, one rounding
    float va, vb;
    ss_unpack(one_m, va, vb);
    const ss_f2 v = ss_pack(fmaxf(va, 0.0f), fmaxf(vb, 0.0f));
    const ss_f2 v2 = ss_mul2(v, v);
    const ss_f2 v3 = ss_mul2(v2, v);
    const ss_f2 outer = ss_mul2(v3, ss_pack(P.a_s2, P.a_s2));
    ss_f2 inner = ss_fma2(v, ss_pack(-P.a_s6, -P.a_s6), ss_pack(P.a_sigma, P.a_sigma));  // fnmadd(v, 6 sigma, sigma)
    inner = ss_fma2(v2, ss_pack(P.a_s12, P.a_s12), inner);
    inner = ss_fma2(v3, ss_pack(-P.a_s6, -P.a_s6), inner);
    float ia, ib, oa, ob;
    ss_unpack(inner, ia, ib);
    ss_unpack(outer, oa, ob);
    return ss_pack(qa <= 0.5f ? ia : oa, qb <= 0.5f ? ib : ob);
}

template <bool GLOBAL, bool COUNT, int CAP, int LIST, typename SLICE>
__device__ __forceinline__ void ss_exact_brick(const SsDev &P, const SsXwArgs &A, SLICE &S, const uint32_t work, const int lane) {
    if (lane == 0) { ss_mbar_init(&S.mbar, 1); ss_mbar_fence_init(); }
    const int nb = P.nb;
    const uint32_t brick_lin = A.bricks[work];
    int bx, by, bz, tile_idx;
    {
        uint32_t q = brick_lin;
        bz = (int)(q % (uint32_t)nb); q /= (uint32_t)nb;
        by = (int)(q % (uint32_t)nb); q /= (uint32_t)nb;
        bx = (int)(q % (uint32_t)nb); tile_idx = (int)(q / (uint32_t)nb);
    }
    // per-box flags of the brick: 16 bytes; sub-box (a, b, c) needs values when either of its two 2x4x4 boxes is flagged
    uint32_t need8 = 0;
    {
        const uint8_t f = lane < SS_LS_WARPS ? A.wflag[(size_t)brick_lin * SS_LS_WARPS + lane] : (uint8_t)0;
        const uint32_t fl = __ballot_sync(0xffffffffu, f != 0);          // bit (i2 * 4 + hb * 2 + hc), i2 = plane pair 0..3
        for (int b = 0; b < 8; ++b) {
            const int ha = (b >> 2) & 1, hbc = b & 3;
            if (fl & ((1u << ((2 * ha) * 4 + hbc)) | (1u << ((2 * ha + 1) * 4 + hbc)))) need8 |= 1u << b;
        }
    }
    if (!need8) return;
    const SsTile T = A.tile_tab[tile_idx];
    const bool sparse = GLOBAL || T.sparse != 0;

    // ---- candidate runs (FIX mode: no extension planes), one per lane
    const int2 rx = A.brick_rng[bx], ry = A.brick_rng[by], rz = A.brick_rng[bz];
    const int nyr = ry.y - ry.x + 1;
    const int nruns = (rx.y - rx.x + 1) * nyr;                  // host guarantees <= 32 for this variant
    uint32_t run_a = 0, run_len = 0;
    if (lane < nruns) {
        const int X = rx.x + lane / nyr, Y = ry.x + lane % nyr;
        uint32_t a = 0xffffffffu, b = 0;
        const uint32_t base = T.s * (uint32_t)P.nbin_sub + (uint32_t)((X * P.nbin + Y) * P.nbin);
        for (int Z = rz.x; Z <= rz.y; ++Z) {
            const uint32_t st = A.bin_start[base + Z];
            if (st != 0xffffffffu) { if (a == 0xffffffffu) a = st; b = A.bin_end[base + Z]; }
        }
        if (a != 0xffffffffu) { run_a = a; run_len = b - a; }
    }
    uint32_t incl = run_len;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
    const uint32_t run_dst = incl - run_len;
    const int C = (int)__shfl_sync(0xffffffffu, incl, 31);
    const int pmax = P.np - 1;
    float *const tile = A.tiles + (size_t)tile_idx * P.np * P.np * P.np;
    const int li = lane >> 3, lj = (lane >> 1) & 3, lk = (lane & 1) * 2;

    if (C == 0) {
        // nothing in reach: the exact value is 0 (the tile may hold markers from an earlier pass only if C > 0, so this is
        // defensive): store zeros for the requested sub-boxes
        for (int b = 0; b < 8; ++b) {
            if (!(need8 & (1u << b))) continue;
            const int i = bx * 8 + 4 * ((b >> 2) & 1) + li, j = by * 8 + 4 * ((b >> 1) & 1) + lj, k = bz * 8 + 4 * (b & 1) + lk;
            if (i <= pmax && j <= pmax) { if (k <= pmax) tile[((size_t)i * P.np + j) * P.np + k] = 0.0f; if (k + 1 <= pmax) tile[((size_t)i * P.np + j) * P.np + k + 1] = 0.0f; }
        }
        return;
    }
    if (C > CAP) {
        if (lane == 0) A.fallback[1 + atomicAdd(&A.fallback[0], 1u)] = brick_lin;
        return;
    }

    // ---- stage: records by bulk copies; sort keys and remainder splits by coalesced loads, run by run
    __syncwarp();
    if (lane == 0) ss_mbar_arrive_expect_tx(&S.mbar, (uint32_t)C * 16u);
    __syncwarp();
    if (run_len) ss_bulk_g2s(&S.rec[run_dst], A.rec + run_a, run_len * 16u, &S.mbar);
    for (int r = 0; r < nruns; ++r) {
        const uint32_t a = __shfl_sync(0xffffffffu, run_a, r), len = __shfl_sync(0xffffffffu, run_len, r), dst = __shfl_sync(0xffffffffu, run_dst, r);
        for (uint32_t t = lane; t < len; t += 32) {
            S.key[dst + t] = ((unsigned long long)A.pidx[a + t] << 32) | (unsigned long long)(dst + t);
            S.ks[dst + t] = (uint16_t)max(A.ksplit[a + t], 0);
        }
    }
    int npow = 32;
    while (npow < C) npow <<= 1;
    for (int t = C + lane; t < npow; t += 32) S.key[t] = ~0ull;
    __syncwarp();
    ss_warp_bitonic(S.key, npow, lane);
    ss_mbar_wait(&S.mbar, 0);
    __syncwarp();

    const float cull2 = (GLOBAL ? P.rev2 : (sparse ? P.h2m : P.h2)) * 1.0001f;
    unsigned hits = 0;
    for (int b = 0; b < 8; ++b) {
        if (!(need8 & (1u << b))) continue;
        const int ha = (b >> 2) & 1, hb = (b >> 1) & 1, hc = b & 1;
        const int i0 = bx * 8 + 4 * ha, j0 = by * 8 + 4 * hb, k0 = bz * 8 + 4 * hc;
        if (i0 > pmax || j0 > pmax || k0 > pmax) continue;
        const int i = i0 + li, j = j0 + lj, k = k0 + lk;
        const bool row_ok = i <= pmax && j <= pmax;
        const bool vA = row_ok && k <= pmax, vB = row_ok && k + 1 <= pmax;
        const int gi = T.gbase[0] + i, gj = T.gbase[1] + j, gk = T.gbase[2] + k;
        // grid point coordinates from GLOBAL indices (dense_subdomains.rs:1068, :1101-1102; scalar: uniform_grid.rs:418-425)
        const float gx = __fadd_rn(__fmul_rn((float)gi, P.c), P.gmin[0]);
        const float gy = __fadd_rn(__fmul_rn((float)gj, P.c), P.gmin[1]);
        const float gzA = sparse ? __fadd_rn(P.gmin[2], __fmul_rn((float)gk, P.c)) : __fmaf_rn((float)gk, P.c, P.gmin[2]);
        const float gzB = sparse ? __fadd_rn(P.gmin[2], __fmul_rn((float)(gk + 1), P.c)) : __fmaf_rn((float)(gk + 1), P.c, P.gmin[2]);
        // sub-box in world coordinates (culling only)
        const int i1 = min(i0 + 3, pmax), j1 = min(j0 + 3, pmax), k1 = min(k0 + 3, pmax);
        const float bxl = fmaf((float)(T.gbase[0] + i0), P.c, P.gmin[0]), bxh = fmaf((float)(T.gbase[0] + i1), P.c, P.gmin[0]);
        const float byl = fmaf((float)(T.gbase[1] + j0), P.c, P.gmin[1]), byh = fmaf((float)(T.gbase[1] + j1), P.c, P.gmin[1]);
        const float bzl = fmaf((float)(T.gbase[2] + k0), P.c, P.gmin[2]), bzh = fmaf((float)(T.gbase[2] + k1), P.c, P.gmin[2]);
        // ---- candidates within the support of the sub-box, in ascending particle index
        int nlist = 0;
        bool overflow = false;
        for (int w = 0; w < ((C + 31) >> 5); ++w) {
            const int rnk = w * 32 + lane;
            bool keep = false;
            int slot = 0;
            if (rnk < C) {
                slot = (int)(S.key[rnk] & 0xffffu);
                const float4 r = S.rec[slot];
                keep = ss_axis_d2(r.x, bxl, bxh) + ss_axis_d2(r.y, byl, byh) + ss_axis_d2(r.z, bzl, bzh) < cull2;
            }
            const uint32_t mword = __ballot_sync(0xffffffffu, keep);
            const int pos = nlist + __popc(mword & ((1u << lane) - 1u));
            if (keep && pos < LIST) S.list[pos] = (uint16_t)slot;
            nlist += __popc(mword);
        }
        if (nlist > LIST) overflow = true;
        __syncwarp();
        if (overflow) {
            // pathological clustering: leave the whole brick to k_levelset (its flags are untouched; values written so far are exact)
            if (lane == 0) A.fallback[1 + atomicAdd(&A.fallback[0], 1u)] = brick_lin;
            return;
        }
        float phiA = 0.0f, phiB = 0.0f;
        if (GLOBAL) {
            for (int n = 0; n < nlist; ++n) {
                const float4 r = S.rec[S.list[n]];
                int im[3]; float d0[3];
                if (!ss_global_candidate(P, r, im, d0)) continue;                      // skipped particle (density_map.rs:645-656)
                ss_accumulate_global(P, r, im, d0, gi, gj, gk, phiA, hits);
                ss_accumulate_global(P, r, im, d0, gi, gj, gk + 1, phiB, hits);
            }
        } else if (sparse) {
            for (int n = 0; n < nlist; ++n) {
                const float4 r = S.rec[S.list[n]];
                ss_accumulate<true, false>(P, r, 0, k, gx, gy, gzA, phiA, hits);
                ss_accumulate<true, false>(P, r, 0, k + 1, gx, gy, gzB, phiB, hits);
            }
        } else {
            const ss_f2 ngz = ss_pack(-gzA, -gzB);
            ss_f2 phi = ss_pack(0.0f, 0.0f);
            for (int n = 0; n < nlist; ++n) {
                const int slot = S.list[n];
                const float4 r = S.rec[slot];
                const int ks = (int)S.ks[slot];
                const float dx = __fsub_rn(r.x, gx), dy = __fsub_rn(r.y, gy);
                const float t = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
                const ss_f2 dz = ss_add2(ss_pack(r.z, r.z), ngz);                      // r.z - gz, exact negation
                const ss_f2 d2 = ss_fma2(dz, dz, ss_pack(t, t));                       // fma(dz, dz, fma(dx, dx, dy * dy))
                float d2a, d2b;
                ss_unpack(d2, d2a, d2b);
                const bool inA = d2a < P.h2, inB = d2b < P.h2;
                if (!__any_sync(0xffffffffu, inA || inB)) continue;
                const ss_f2 wk = ss_kernel_avx2(P, __fsqrt_rn(d2a), __fsqrt_rn(d2b));
                float wa, wb;
                ss_unpack(wk, wa, wb);
                wa = inA ? wa : 0.0f; wb = inB ? wb : 0.0f;                            // outside the support: exact zero term
                if (COUNT) hits += (inA ? 1u : 0u) + (inB ? 1u : 0u);
                if (ks > k1) {
                    phi = ss_fma2(ss_pack(wa, wb), ss_pack(r.w, r.w), phi);            // fmadd lanes (dense_subdomains.rs:1110-1114)
                } else {
                    float pa, pb;
                    ss_unpack(phi, pa, pb);
                    pa = (k < ks) ? __fmaf_rn(wa, r.w, pa) : __fadd_rn(pa, __fmul_rn(wa, r.w));         // remainder lanes (:1117-1128)
                    pb = (k + 1 < ks) ? __fmaf_rn(wb, r.w, pb) : __fadd_rn(pb, __fmul_rn(wb, r.w));
                    phi = ss_pack(pa, pb);
                }
            }
            ss_unpack(phi, phiA, phiB);
        }
        if (vA) tile[((size_t)i * P.np + j) * P.np + k] = phiA;
        if (vB) tile[((size_t)i * P.np + j) * P.np + k + 1] = phiB;
        __syncwarp();                                           // the list is rebuilt for the next sub-box
    }
    if (COUNT && A.pairs) {
        unsigned long long np_ = hits;
        for (int o = 16; o > 0; o >>= 1) np_ += __shfl_xor_sync(0xffffffffu, np_, o);
        if (lane == 0 && np_) atomicAdd(A.pairs, np_);
    }
}

template <bool GLOBAL, bool COUNT>
__global__ void __launch_bounds__(SS_XW_THREADS, 5)
k_exact_warp(SsDev P, SsXwArgs A) {
    __shared__ SsXwSlice s_slice[SS_XW_WARPS];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t work = blockIdx.x * SS_XW_WARPS + wib;
    if (work >= A.n_bricks) return;
    ss_exact_brick<GLOBAL, COUNT, SS_XW_CAP, SS_XW_LIST>(P, A, s_slice[wib], work, lane);
}

// one warp per CTA: bricks with up to 1024 candidates
template <bool GLOBAL, bool COUNT>
__global__ void __launch_bounds__(32, 8)
k_exact_warp_mid(SsDev P, SsXwArgs A) {
    __shared__ SsXwSliceMid s_slice;
    if (blockIdx.x >= A.n_bricks) return;
    ss_exact_brick<GLOBAL, COUNT, SS_XW_MID_CAP, SS_XW_MID_LIST>(P, A, s_slice, blockIdx.x, (int)(threadIdx.x & 31));
}

// one warp per CTA, slice in dynamic shared memory (sizeof(SsXwSliceBig) bytes): bricks with up to 4096 candidates
template <bool GLOBAL, bool COUNT>
__global__ void __launch_bounds__(32, 2)
k_exact_warp_big(SsDev P, SsXwArgs A) {
    if (blockIdx.x >= A.n_bricks) return;
    ss_exact_brick<GLOBAL, COUNT, SS_XW_BIG_CAP, SS_XW_BIG_LIST>(P, A, *reinterpret_cast<SsXwSliceBig *>(ss_dyn_smem), blockIdx.x, (int)(threadIdx.x & 31));
}
