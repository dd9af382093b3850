// ss_mc.cuh -- brick passes of the subdomain path, one WARP per 8x8x8-point brick (sm_100a): marching cubes in two launches
// (count, emit) and the marker fix-up sweep.  They replace the CTA-per-brick passes of ss_kernels.cuh (k_mc_count / k_mc_verts /
// k_mc_tris / k_fixup_flags: 512 threads and two CTA barriers per brick, ~1.5 us per brick in flight, three launches over
// 1.9 M bricks at 50 M particles), which stay in use for the global (non-decomposed) path.
//
// Reference: per-cell marching cubes of a subdomain (dense_subdomains.rs:1470-1568): a vertex on every grid edge whose endpoints
// lie on different sides (value > threshold), triangles from the 256-case table with the reversed triplet order.  Values and
// classifications are the level-set stage's; everything here is integer work plus the vertex interpolation, which keeps the
// reference's operation sequence (`a (1 - alpha) + b alpha`, :1516-1519).
//
// Data flow per brick (a warp; no CTA barrier, no shared-memory tile of values):
//   * the "above" bits of the brick's 9x9x9 point neighbourhood are gathered as 81 ROW MASKS (9 bits along z): lane = row, nine
//     loads each; every cell's case index and every point's edge mask are then bit operations on four row masks;
//   * lane = two (i, j) columns of the brick; the 8 points / cells of a column are walked in k;
//   * count pass: per-point edge masks (vmask), the LOCAL vertex offset of every point that owns a vertex (voff), brick totals;
//   * after the exclusive scans of the brick totals, the emit pass writes vertices, edge keys, the boundary list and the
//     triangles in one launch: the global id of a vertex owned by a point of a neighbouring brick is
//         vbase + vblk_off[list position of that brick] + voff[point] + (number of lower-axis vertices of the point),
//     so no pass has to wait for the vertex ids of another brick.
// Vertex / triangle order inside a brick differs from the CTA passes (column-major per lane); the mesh is the same set.
#pragma once

#define SS_MW_WARPS 8                  // bricks in flight per CTA

struct SsMwSlice {
    uint16_t rm[128];                  // row masks (up to 10 x 10 rows for the fix-up sweep; 9 x 9 for marching cubes)
    uint16_t ro[128];                  // fix-up sweep: "outside" bits of the same rows
    uint32_t base[8];                  // emit pass: vertex id base of the own brick [0] and of the 7 (+x, +y, +z) neighbours
};

// ---- row masks: bit c of rm[a * 9 + b] = value of tile point (o + (a, b, c)) > thr, for a, b, c in 0..8; points outside the tile: 0
__device__ __forceinline__ void ss_mw_rows(const SsDev &P, const float *__restrict__ phi, int ox, int oy, int oz, SsMwSlice &S, int lane) {
    const int np = P.np;
    for (int r = lane; r < 81; r += 32) {
        const int a = r / 9, b = r - 9 * a;
        const int gi = ox + a, gj = oy + b;
        uint32_t m = 0;
        if (gi < np && gj < np) {
            const float *row = phi + ((size_t)gi * np + gj) * np + oz;
            const int nc = min(9, np - oz);
#pragma unroll
            for (int c = 0; c < 9; ++c) if (c < nc && row[c] > P.thr) m |= 1u << c;
        }
        S.rm[r] = (uint16_t)m;
    }
    __syncwarp();
}

// per (i, j) column of the brick: edge masks of its 8 points and the mask of its non-trivial cells
struct SsMwColumn {
    uint32_t ex, ey, ez;               // bit k: point k owns a vertex on its +x / +y / +z edge
    uint32_t cells;                    // bit k: cell k has corners on both sides (can emit triangles)
    uint32_t m00, m10, m01, m11;       // row masks of the four z-rows of the column
};
__device__ __forceinline__ SsMwColumn ss_mw_column(const SsDev &P, const SsMwSlice &S, int a, int b, int i, int j, int oz) {
    SsMwColumn C;
    const int np = P.np;
    C.ex = C.ey = C.ez = C.cells = 0; C.m00 = C.m10 = C.m01 = C.m11 = 0;
    if (i >= np || j >= np) return C;
    C.m00 = S.rm[a * 9 + b]; C.m10 = S.rm[(a + 1) * 9 + b]; C.m01 = S.rm[a * 9 + b + 1]; C.m11 = S.rm[(a + 1) * 9 + b + 1];
    const uint32_t pts = (1u << min(8, np - oz)) - 1u;                      // own points of the column inside the tile
    const uint32_t zed = (1u << max(0, min(8, np - 1 - oz))) - 1u;          // ... whose +z neighbour is inside the tile too
    if (i + 1 < np) C.ex = (C.m00 ^ C.m10) & pts;
    if (j + 1 < np) C.ey = (C.m00 ^ C.m01) & pts;
    C.ez = (C.m00 ^ (C.m00 >> 1)) & zed;
    if (i < P.S && j < P.S) {
        const uint32_t any = C.m00 | C.m10 | C.m01 | C.m11, all = C.m00 & C.m10 & C.m01 & C.m11;
        const uint32_t triv = (~any & (~any >> 1)) | (all & (all >> 1));    // all eight corners below / all above
        C.cells = ~triv & ((1u << max(0, min(8, P.S - oz))) - 1u);
    }
    return C;
}
__device__ __forceinline__ int ss_mw_case(const SsMwColumn &C, int k) {
    return (int)(((C.m00 >> k) & 1u) | (((C.m10 >> k) & 1u) << 1) | (((C.m11 >> k) & 1u) << 2) | (((C.m01 >> k) & 1u) << 3) |
                 (((C.m00 >> (k + 1)) & 1u) << 4) | (((C.m10 >> (k + 1)) & 1u) << 5) | (((C.m11 >> (k + 1)) & 1u) << 6) | (((C.m01 >> (k + 1)) & 1u) << 7));
}
__device__ __forceinline__ uint32_t ss_mw_column_tris(const SsMwColumn &C) {
    uint32_t nt = 0;
    for (uint32_t m = C.cells; m; m &= m - 1u) nt += c_num_tris[ss_mw_case(C, __ffs(m) - 1)];
    return nt;
}
// exclusive prefix over the lanes of a (vertices | triangles << 16) pair count; total in `total`
__device__ __forceinline__ uint32_t ss_mw_scan(uint32_t packed, int lane, uint32_t &total) {
    uint32_t incl = packed;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - packed;
}

// Pass 1 over the bricks k_brick_classify listed: edge masks, local vertex offsets, brick totals.
__global__ void __launch_bounds__(SS_MW_WARPS * 32)
k_mc_count_warp(SsDev P, const float *__restrict__ tiles, const uint32_t *__restrict__ list, uint32_t n_list, uint8_t *__restrict__ vmask,
                uint32_t *__restrict__ voff, uint32_t *__restrict__ vblk, uint32_t *__restrict__ tblk) {
    __shared__ SsMwSlice s_slice[SS_MW_WARPS];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * SS_MW_WARPS + wib;
    if (w >= n_list) return;
    SsMwSlice &S = s_slice[wib];
    const SsBrick B = ss_brick_from_lin(P, list[w]);
    const int np = P.np, ox = B.bx * 8, oy = B.by * 8, oz = B.bz * 8;
    const size_t tbase = (size_t)B.tile * np * np * np;
    ss_mw_rows(P, tiles + tbase, ox, oy, oz, S, lane);
    SsMwColumn C[2];
    uint32_t nv[2], nt[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int col = lane + 32 * h, a = col >> 3, b = col & 7;
        C[h] = ss_mw_column(P, S, a, b, ox + a, oy + b, oz);
        nv[h] = __popc(C[h].ex) + __popc(C[h].ey) + __popc(C[h].ez);
        nt[h] = ss_mw_column_tris(C[h]);
    }
    uint32_t total;
    uint32_t pre = ss_mw_scan((nv[0] + nv[1]) | ((nt[0] + nt[1]) << 16), lane, total) & 0xffffu;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int col = lane + 32 * h, a = col >> 3, b = col & 7;
        const size_t pt0 = tbase + ((size_t)(ox + a) * np + (oy + b)) * np + oz;
        for (uint32_t m = C[h].ex | C[h].ey | C[h].ez; m; m &= m - 1u) {
            const int k = __ffs(m) - 1;
            const uint32_t mask = ((C[h].ex >> k) & 1u) | (((C[h].ey >> k) & 1u) << 1) | (((C[h].ez >> k) & 1u) << 2);
            vmask[pt0 + k] = (uint8_t)mask;
            voff[pt0 + k] = pre;
            pre += __popc(mask);
        }
    }
    if (lane == 0) { vblk[w] = total & 0xffffu; tblk[w] = total >> 16; }
}

// Pass 2: vertices (dense_subdomains.rs:1498-1538), edge keys, boundary list and triangles (:1470-1552) of every listed brick.
__global__ void __launch_bounds__(SS_MW_WARPS * 32)
k_mc_emit_warp(SsDev P, const float *__restrict__ tiles, const uint32_t *__restrict__ list, uint32_t n_list, const uint8_t *__restrict__ vmask,
               const uint32_t *__restrict__ voff, const uint32_t *__restrict__ vblk, const uint32_t *__restrict__ tblk,
               const uint32_t *__restrict__ vblk_off, const uint32_t *__restrict__ tblk_off, const uint32_t *__restrict__ flag_mc,
               const uint32_t *__restrict__ off_mc, const SsTile *__restrict__ tile_tab, SsMcOut O) {
    __shared__ SsMwSlice s_slice[SS_MW_WARPS];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * SS_MW_WARPS + wib;
    if (w >= n_list) return;
    const uint32_t nvb = vblk[w], ntb = tblk[w];
    if (nvb == 0 && ntb == 0) return;
    SsMwSlice &S = s_slice[wib];
    const SsBrick B = ss_brick_from_lin(P, list[w]);
    const int np = P.np, nb = P.nb, ox = B.bx * 8, oy = B.by * 8, oz = B.bz * 8;
    const size_t tbase = (size_t)B.tile * np * np * np;
    const float *phi = tiles + tbase;
    // vertex id bases: own brick and the seven neighbours a cell of this brick can reach (q = fx | fy << 1 | fz << 2)
    if (lane < 8) {
        const int x = B.bx + (lane & 1), y = B.by + ((lane >> 1) & 1), z = B.bz + (lane >> 2);
        uint32_t base = 0;
        if (x < nb && y < nb && z < nb) {
            const uint32_t lin = (uint32_t)((((size_t)B.tile * nb + x) * nb + y) * nb + z);
            if (flag_mc[lin]) base = O.vbase + vblk_off[off_mc[lin]];
        }
        S.base[lane] = base;
    }
    ss_mw_rows(P, phi, ox, oy, oz, S, lane);       // ends with __syncwarp: bases and row masks visible
    SsMwColumn C[2];
    uint32_t nt[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int col = lane + 32 * h, a = col >> 3, b = col & 7;
        C[h] = ss_mw_column(P, S, a, b, ox + a, oy + b, oz);
        nt[h] = ss_mw_column_tris(C[h]);
    }
    // ---- vertices of the own points
    if (nvb) {
        const SsTile T = tile_tab[B.tile];
        const float thr = P.thr;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int col = lane + 32 * h, a = col >> 3, b = col & 7;
            const int i = ox + a, j = oy + b;
            for (uint32_t m = C[h].ex | C[h].ey | C[h].ez; m; m &= m - 1u) {
                const int kk = __ffs(m) - 1, k = oz + kk;
                const int l = (i * np + j) * np + k;
                uint32_t vid = S.base[0] + voff[tbase + l];
                const float av = phi[l];
                const int o[3] = { i, j, k };
                float oc[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) oc[d] = ss_coord(T.smin[d], o[d], P.c);
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    const uint32_t em = ax == 0 ? C[h].ex : (ax == 1 ? C[h].ey : C[h].ez);
                    if (!((em >> kk) & 1u)) continue;
                    const int stride = ax == 0 ? np * np : (ax == 1 ? np : 1);
                    const float bval = phi[l + stride];
                    const float alpha = __fdiv_rn(__fsub_rn(thr, av), __fsub_rn(bval, av));      // dense_subdomains.rs:1516-1519
                    const float one_m = __fsub_rn(1.0f, alpha);
                    float pos[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float tc = (d == ax) ? ss_coord(T.smin[d], o[d] + 1, P.c) : oc[d];
                        pos[d] = __fadd_rn(__fmul_rn(oc[d], one_m), __fmul_rn(tc, alpha));
                    }
                    O.verts[3 * (size_t)vid] = pos[0]; O.verts[3 * (size_t)vid + 1] = pos[1]; O.verts[3 * (size_t)vid + 2] = pos[2];
                    const unsigned long long key = ss_edge_key(T.gbase[0] + i, T.gbase[1] + j, T.gbase[2] + k, ax);
                    O.vkeys[vid] = key;
                    // boundary edge: an orthogonal coordinate on a tile face (uniform_grid.rs:332-338)
                    bool boundary = false;
#pragma unroll
                    for (int d = 0; d < 3; ++d) if (d != ax && (o[d] == 0 || o[d] == P.S)) boundary = true;
                    if (boundary) {
                        const uint32_t slot = atomicAdd(O.bcount, 1u);
                        if (slot < O.bcap) { O.bkeys[slot] = key; O.bids[slot] = vid; }
                    }
                    ++vid;
                }
            }
        }
    }
    // ---- triangles of the own cells
    if (ntb) {
        uint32_t total;
        uint32_t tid = O.tbase + tblk_off[w] + (ss_mw_scan((nt[0] + nt[1]) << 16, lane, total) >> 16);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int col = lane + 32 * h, a = col >> 3, b = col & 7;
            const int i = ox + a, j = oy + b;
            for (uint32_t m = C[h].cells; m; m &= m - 1u) {
                const int kk = __ffs(m) - 1;
                const int idx = ss_mw_case(C[h], kk);
                const int ntc = c_num_tris[idx];
                const int l = (i * np + j) * np + oz + kk;
                for (int q = 0; q < ntc; ++q) {
#pragma unroll
                    for (int mth = 0; mth < 3; ++mth) {
                        const int le = c_tri_table[idx][3 * q + (2 - mth)];              // reversed triplet, lut.rs:338-342
                        const int ax = c_edge_axis[le];
                        const int da = c_edge_org[le][0], db = c_edge_org[le][1], dc = c_edge_org[le][2];
                        const int lo = l + da * np * np + db * np + dc;
                        const int qn = ((a + da) >> 3) | (((b + db) >> 3) << 1) | (((kk + dc) >> 3) << 2);
                        const uint32_t below = vmask[tbase + lo] & ((1u << ax) - 1u);
                        O.tris[3 * (size_t)tid + mth] = S.base[qn] + voff[tbase + lo] + __popc(below);
                    }
                    ++tid;
                }
            }
        }
    }
}

// Marker points (certified inside, value unknown) that touch an outside point along a grid edge carry a surface-crossing edge,
// so they need their exact value: flags their 2x4x4 box and lists their brick for the exact pass (same result as
// k_fixup_flags).  Rows here are 10 long: the brick's points plus one neighbour on either side, along every axis.
__global__ void __launch_bounds__(SS_MW_WARPS * 32)
k_fixup_flags_warp(SsDev P, const float *__restrict__ tiles, const uint32_t *__restrict__ list, uint32_t n_list, uint8_t *__restrict__ wflag,
                   uint32_t *__restrict__ fix_bricks, uint32_t *__restrict__ nfix /* [0]: bricks, [1]: points */) {
    __shared__ SsMwSlice s_slice[SS_MW_WARPS];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * SS_MW_WARPS + wib;
    if (w >= n_list) return;
    SsMwSlice &S = s_slice[wib];
    const SsBrick B = ss_brick_from_lin(P, list[w]);
    const int np = P.np, ox = B.bx * 8 - 1, oy = B.by * 8 - 1, oz = B.bz * 8 - 1;
    const float *phi = tiles + (size_t)B.tile * np * np * np;
    // row (a, b), a, b in 0..9 <-> tile point (ox + a, oy + b, oz + c), c in 0..9: marker bits and outside bits
    for (int r = lane; r < 100; r += 32) {
        const int a = r / 10, b = r - 10 * a;
        const int gi = ox + a, gj = oy + b;
        uint32_t mk = 0, out = 0;
        if (gi >= 0 && gj >= 0 && gi < np && gj < np) {
            const float *row = phi + ((size_t)gi * np + gj) * np;
#pragma unroll
            for (int c = 0; c < 10; ++c) {
                const int gk = oz + c;
                if (gk >= 0 && gk < np) {
                    const float v = row[gk];
                    if (v == SS_MARKER) mk |= 1u << c;
                    if (!(v > P.thr)) out |= 1u << c;
                }
            }
        }
        S.rm[r] = (uint16_t)mk; S.ro[r] = (uint16_t)out;
    }
    __syncwarp();
    uint32_t npts = 0;
    bool any = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int col = lane + 32 * h, a = (col >> 3) + 1, b = (col & 7) + 1;          // own point column in row coordinates
        const uint32_t o0 = S.ro[a * 10 + b];
        const uint32_t nbr = S.ro[(a - 1) * 10 + b] | S.ro[(a + 1) * 10 + b] | S.ro[a * 10 + b - 1] | S.ro[a * 10 + b + 1] | (o0 << 1) | (o0 >> 1);
        const uint32_t hit = S.rm[a * 10 + b] & nbr & 0x1feu;                              // own points: bits 1..8
        if (hit) {
            const int box = (((a - 1) >> 1) << 2) | (((b - 1) >> 2) << 1);
            if (hit & 0x01eu) wflag[(size_t)B.lin * SS_LS_WARPS + box] = 1;
            if (hit & 0x1e0u) wflag[(size_t)B.lin * SS_LS_WARPS + box + 1] = 1;
            npts += __popc(hit);
            any = true;
        }
    }
    for (int o = 16; o > 0; o >>= 1) npts += __shfl_xor_sync(0xffffffffu, npts, o);
    const bool brick_hit = __any_sync(0xffffffffu, any);
    if (lane == 0 && brick_hit) {
        fix_bricks[atomicAdd(&nfix[0], 1u)] = B.lin;
        atomicAdd(&nfix[1], npts);
    }
}
