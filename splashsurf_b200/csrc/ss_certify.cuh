// ss_certify.cuh -- level-set variant 1: certification in its own, barrier-light kernel (candidate for round 2; selected with
// ss_context_set_levelset_variant(ctx, 1), the default stays k_levelset's fused certify + exact pass).
//
// Why: in k_levelset 27 % of the warp-time is spent at CTA barriers (profiles/r1d_levelset_source_breakdown.txt): every warp
// waits for warp 0's run scan and, after certifying its own 32 points, for the slowest warp of the brick before it may
// store its markers -- although 94 % of the bricks of a bulk fluid never enter the exact path.  Here a brick's warps
//   * compute the candidate-run table redundantly in registers (shuffles, no shared table, no scan barrier),
//   * stage only the particle records (x, y, z, V) -- no sort keys, no k-split -- behind ONE barrier,
//   * certify their box, store markers or raise a per-box flag, and retire independently.
// Boxes whose certification fails are then evaluated by k_levelset in SS_LS_FIX mode (an extra, small launch over the flagged
// bricks), exactly like the boxes the fix-up sweep flags.  Results are identical to the fused pass by construction: the same
// ss_certify_box decides, the same SS_LS_FIX code evaluates.
//
// Per-box state wstate[brick * 16 + box] : 0 untouched (all zero), 1 every valid point carries SS_MARKER, 2 needs exact values,
// 3 exact values already in place (no candidate within the support of any point of the box: the pre-zeroed tile IS the result).
#pragma once

struct SsCertArgs {
    const uint32_t *bin_start, *bin_end;
    const float4 *rec;
    const SsTile *tile_tab;
    const int2 *brick_rng;
    const uint4 *work_desc;      // per listed brick: (linear brick index, tile, bx | by << 10 | bz << 20, 0)
    float *tiles;
    uint8_t *wstate;             // [batch][nb^3][16]
};

// work list with decoded brick coordinates (no runtime divisions in the kernel)
__global__ void k_compact_desc(SsDev P, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ off, uint32_t n, uint4 *__restrict__ desc) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n || !flag[b]) return;
    const uint32_t nb = (uint32_t)P.nb;
    uint32_t q = b;
    const uint32_t bz = q % nb; q /= nb;
    const uint32_t by = q % nb; q /= nb;
    const uint32_t bx = q % nb; const uint32_t tile = q / nb;
    desc[off[b]] = make_uint4(b, tile, bx | (by << 10) | (bz << 20), 0u);
}

// marks the standard 2x4x4 boxes of brick (vbx, vby, vbz) that the point box W overlaps
__device__ __forceinline__ void ss_mark_boxes(const SsDev &P, uint8_t *__restrict__ wstate, int tile, int vbx, int vby, int vbz, const SsWarpBox &W, uint8_t value) {
    const size_t base = ((((size_t)tile * P.nb + vbx) * P.nb + vby) * P.nb + vbz) * SS_LS_WARPS;
    const int i0 = W.i0 - 8 * vbx, j0 = W.j0 - 8 * vby, k0 = W.k0 - 8 * vbz;
    for (int a = i0 >> 1; a <= min((i0 + W.dx - 1) >> 1, 3); ++a)
        for (int b = j0 >> 2; b <= min((j0 + W.dy - 1) >> 2, 1); ++b)
            for (int c = k0 >> 2; c <= min((k0 + W.dz - 1) >> 2, 1); ++c) wstate[base + (size_t)(a * 4 + b * 2 + c)] = value;
}

// true when some staged candidate lies within the kernel support of the warp's point box (the cull of ss_exact_box)
template <bool GLOBAL>
__device__ __forceinline__ bool ss_box_has_candidate(const SsDev &P, const SsLanePoint &L, bool sparse, const float4 *s_rec, int C, int lane) {
    const float cull2 = (GLOBAL ? P.rev2 : (sparse ? P.h2m : P.h2)) * 1.0001f;
    for (int w = 0; w < ((C + 31) >> 5); ++w) {
        const int c = w * 32 + lane;
        if (__any_sync(0xffffffffu, c < C && ss_box_dist2(L, s_rec[c < C ? c : 0]) < cull2)) return true;
    }
    return false;
}

template <bool GLOBAL>
__global__ void __launch_bounds__(SS_LS_THREADS, 3)
k_certify(SsDev P, SsCertArgs A) {
    __shared__ float4 s_rec[SS_LS_CAP];
    const int nb = P.nb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint4 desc = A.work_desc[blockIdx.x];
    const int tile_idx = (int)desc.y;
    const int bx = (int)(desc.z & 1023u), by = (int)((desc.z >> 10) & 1023u), bz = (int)(desc.z >> 20);
    const bool ext = P.ext_bricks != 0;
    const bool ex = ext && bx == nb - 2, ey = ext && by == nb - 2, ez = ext && bz == nb - 2;
    const SsTile T = A.tile_tab[tile_idx];

    // ---- candidate runs (one per lane, host guarantees <= 32 for this variant) and their exclusive prefix, per warp
    int2 rx = A.brick_rng[bx], ry = A.brick_rng[by], rz = A.brick_rng[bz];
    if (ex) rx.y = A.brick_rng[bx + 1].y;
    if (ey) ry.y = A.brick_rng[by + 1].y;
    if (ez) rz.y = A.brick_rng[bz + 1].y;
    const int nyr = ry.y - ry.x + 1;
    const int nruns = (rx.y - rx.x + 1) * nyr;
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
    if (C == 0) return;                                         // tile is pre-zeroed: phi = 0 exactly (wstate stays 0)

    const SsWarpBox Wm = { bx * 8 + (warp >> 2) * 2, by * 8 + ((warp >> 1) & 1) * 4, bz * 8 + (warp & 1) * 4, 2, 4, 4 };
    SsWarpBox We = Wm;
    int vbx = bx, vby = by, vbz = bz;
    const bool has_ext = (ex || ey || ez) && warp < 10 && ss_ext_task(warp, bx, by, bz, ex, ey, ez, We, vbx, vby, vbz);
    const bool sparse = GLOBAL || T.sparse != 0;

    if (C > SS_LS_CAP) {
        // oversized brick: everything goes to the exact pass (k_levelset's oversized path honours the per-box flags)
        const SsLanePoint L = ss_lane_point(P, T, Wm, tile_idx, lane, sparse);
        if (L.warp_valid && lane == 0) A.wstate[(size_t)desc.x * SS_LS_WARPS + warp] = 2;
        if (has_ext) { const SsLanePoint Le = ss_lane_point(P, T, We, tile_idx, lane, sparse); if (Le.warp_valid && lane == 0) ss_mark_boxes(P, A.wstate, tile_idx, vbx, vby, vbz, We, 2); }
        return;
    }

    // ---- stage the records (bin order): warp w takes runs w, w + 16
    for (int r = warp; r < nruns; r += SS_LS_WARPS) {
        const uint32_t a = __shfl_sync(0xffffffffu, run_a, r), len = __shfl_sync(0xffffffffu, run_len, r), dst = __shfl_sync(0xffffffffu, run_dst, r);
        for (uint32_t t = lane; t < len; t += 32) {
            float4 rc = A.rec[a + t];
            if (GLOBAL) { int im[3]; float d0[3]; if (!ss_global_candidate(P, rc, im, d0)) rc.w = 0.0f; }   // skipped particle: no volume
            s_rec[dst + t] = rc;
        }
    }
    // the candidate sweep starts at the run that holds the brick's own bin
    int w0;
    {
        const int ox = min(max(ss_floor_div(8 * bx, P.be) + P.nlo - rx.x, 0), rx.y - rx.x), oy = min(max(ss_floor_div(8 * by, P.be) + P.nlo - ry.x, 0), nyr - 1);
        w0 = (int)(__shfl_sync(0xffffffffu, run_dst, ox * nyr + oy) >> 5);
        if (w0 >= ((C + 31) >> 5)) w0 = 0;
    }
    __syncthreads();

    {
        const SsLanePoint L = ss_lane_point(P, T, Wm, tile_idx, lane, sparse);
        if (L.warp_valid) {
            const bool ok = ss_certify_box(P, L, s_rec, C, lane, w0);
            if (ok && L.valid) A.tiles[L.out_idx] = SS_MARKER;
            const uint8_t st = ok ? 1 : (ss_box_has_candidate<GLOBAL>(P, L, sparse, s_rec, C, lane) ? 2 : 3);
            if (lane == 0) A.wstate[(size_t)desc.x * SS_LS_WARPS + warp] = st;
        }
    }
    if (has_ext) {
        const SsLanePoint L = ss_lane_point(P, T, We, tile_idx, lane, sparse);
        if (L.warp_valid) {
            const bool ok = ss_certify_box(P, L, s_rec, C, lane, w0);
            if (ok && L.valid) A.tiles[L.out_idx] = SS_MARKER;
            const uint8_t st = ok ? 1 : (ss_box_has_candidate<GLOBAL>(P, L, sparse, s_rec, C, lane) ? 2 : 3);
            if (lane == 0) ss_mark_boxes(P, A.wstate, tile_idx, vbx, vby, vbz, We, st);
        }
    }
}

// per brick: state = max over its boxes (0 untouched, 1 certified, 2 has exact values -- box states 2 and 3), and the flags of
// the exact launch: need[b] = some box of b is in state 2, wflag = per-box "evaluate exactly"
__global__ void k_wstate_reduce(const uint8_t *__restrict__ wstate, uint32_t nbricks, uint8_t *__restrict__ bstate, uint32_t *__restrict__ need,
                                uint8_t *__restrict__ wflag) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbricks) return;
    const uint4 w = *reinterpret_cast<const uint4 *>(wstate + (size_t)b * SS_LS_WARPS);
    const uint32_t q[4] = { w.x, w.y, w.z, w.w };
    uint32_t mx = 0, any_need = 0, f[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        f[k] = 0;
#pragma unroll
        for (int s = 0; s < 32; s += 8) { const uint32_t v = (q[k] >> s) & 0xffu; mx = max(mx, min(v, 2u)); if (v == 2u) { f[k] |= 1u << s; any_need = 1u; } }
    }
    bstate[b] = (uint8_t)mx;
    need[b] = any_need;
    *reinterpret_cast<uint4 *>(wflag + (size_t)b * SS_LS_WARPS) = make_uint4(f[0], f[1], f[2], f[3]);
}
