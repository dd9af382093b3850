// post_emul.cpp -- steps the per-thread post-processing kernels (splashsurf_b200/csrc/ss_post.cuh) and the binning kernels they
// depend on (k_bin_keys, k_mark_starts, k_run_counts, k_records) on the CPU executor (cuda_emul.h).  TEST INFRASTRUCTURE ONLY
// (tests/test_post_emulation.py): it checks the kernels' indexing and arithmetic against the oracle without a GPU; the
// product never links this file.  Sorts and scans (cub on the device) are done by numpy in the test.
#include "cuda_emul.h"
#include "../../splashsurf_b200/csrc/ss_kernels.cuh"
#define SS_POST_KERNELS_ONLY
#include "../../splashsurf_b200/csrc/ss_post.cuh"

template <typename... P, typename... A>
static void run(uint64_t n, unsigned block, void (*kernel)(P...), A... args) {
    emul::launch(dim3((unsigned)((n + block - 1) / block)), dim3(block), kernel, args...);
}

extern "C" {

unsigned emul_sizeof_dev() { return (unsigned)sizeof(SsDev); }

// The SsDev fields the binning / post-processing kernels read; same expressions as run_subdomain_grid, fill_kernel_consts and
// fill_bins in ss_pipeline.cu.
void emul_make_dev(SsDev *D, const float gmin[3], float c, float h, int S, const int nsd[3], float margin) {
    memset(D, 0, sizeof(SsDev));
    for (int d = 0; d < 3; ++d) { D->gmin[d] = gmin[d]; D->nsd[d] = nsd[d]; }
    D->c = c; D->h = h; D->h2 = __fmul_rn(h, h); D->h2m = __fmul_rn(D->h2, 1.01f);
    D->sub_size = __fmul_rn(c, (float)S); D->margin = margin; D->grow = __fmul_rn(margin, 1.5f);
    D->S = S; D->np = S + 1; D->R = (int)ceilf(__fdiv_rn(h, c));
    const float rrr = __fmul_rn(__fmul_rn(h, h), h);
    D->a_hinv = __fdiv_rn(1.0f, h);
    D->a_sigma = __fdiv_rn(8.0f, __fmul_rn(SS_PI_F, rrr));
    D->s_sigma = __fdiv_rn(8.0f, rrr);
    D->s_c_inner = __fdiv_rn(3.0f, __fmul_rn(2.0f, SS_PI_F));
    D->s_c_outer = __fdiv_rn(1.0f, __fmul_rn(4.0f, SS_PI_F));
    D->s_two_thirds = __fdiv_rn(2.0f, 3.0f);
    D->nb = (D->np + 7) / 8;
    D->be = 8 * std::max(1, (7 + 2 * D->R + 39) / 40);
    D->nlo = (D->R + D->be - 1) / D->be;
    D->nbin = ss_floor_div(D->S + D->R, D->be) + D->nlo + 1;
    D->nbin_sub = D->nbin * D->nbin * D->nbin;
    D->inv_c = (float)(1.0 / (double)c);
    D->rr_cells = (float)D->R + 0.01f;
    D->keep_hi = nsd[0];
}
int emul_dev_nbin_sub(const SsDev *D) { return D->nbin_sub; }

// ---- binning (ss_kernels.cuh)
void emul_bin_keys(const SsDev *D, const float *xyz, uint32_t m, const uint32_t *cid, const uint32_t *sub_flat, const uint32_t *pidx,
                   uint32_t *key) {
    run(m, 256, k_bin_keys, *D, xyz, m, cid, sub_flat, pidx, (const uint8_t *)nullptr, key);
}
void emul_bin_tables(const uint32_t *key_sorted, uint32_t m, uint32_t *start, uint32_t *end, uint32_t nkeys) {
    for (uint32_t k = 0; k < nkeys; ++k) start[k] = 0xffffffffu;
    run(m, 256, k_mark_starts, key_sorted, m, start, nkeys);
    run(m, 256, k_run_counts, key_sorted, m, end, nkeys);
}
void emul_records(const SsDev *D, const float *xyz, const float *rho, uint32_t m, const uint32_t *key_sorted, const uint32_t *pidx_sorted,
                  const uint32_t *sub_flat, float *rec /* m x 4 */, int *ksplit) {
    run(m, 256, k_records, *D, xyz, rho, m, key_sorted, pidx_sorted, sub_flat, (float4 *)rec, ksplit);
}

// ---- post-processing (ss_post.cuh)
static SsQuery make_query(const uint32_t *sub_flat, uint32_t nsub, const uint32_t *bin_start, const uint32_t *bin_end, const uint32_t *pidx,
                          const float *rec, const float *rho, float sphere_mass) {
    SsQuery Q{};
    Q.sub_flat = sub_flat; Q.nsub = nsub; Q.bin_start = bin_start; Q.bin_end = bin_end; Q.pidx = pidx; Q.rec = (const float4 *)rec;
    Q.rho = rho; Q.sphere_mass = sphere_mass;
    return Q;
}
#define QUERY_ARGS const uint32_t *sub_flat, uint32_t nsub, const uint32_t *bin_start, const uint32_t *bin_end, const uint32_t *pidx, \
                   const float *rec, const float *rho, float sphere_mass
#define QUERY make_query(sub_flat, nsub, bin_start, bin_end, pidx, rec, rho, sphere_mass)

void emul_weighted_ncount(const SsDev *D, QUERY_ARGS, const uint32_t *key_sorted, uint32_t m, float *wnc) {
    run(m, 128, k_pp_weighted_ncount, *D, QUERY, key_sorted, m, wnc);
}
void emul_interpolate(const SsDev *D, QUERY_ARGS, const float *pts, uint32_t npts, const float *values, int dim, int correction, float *out) {
    if (dim == 1) run(npts, 128, k_pp_interpolate<1>, *D, QUERY, pts, npts, values, correction, out);
    else run(npts, 128, k_pp_interpolate<3>, *D, QUERY, pts, npts, values, correction, out);
}
void emul_sph_normals(const SsDev *D, QUERY_ARGS, const float *pts, uint32_t npts, float *normals) {
    run(npts, 128, k_pp_sph_normals, *D, QUERY, pts, npts, normals);
}
void emul_smoothstep(uint32_t n, const float *wnn, float normalization, float *out) { run(n, 256, k_pp_smoothstep, n, wnn, normalization, out); }

void emul_edge_keys(const uint32_t *tris, uint32_t nt, unsigned long long *keys) { run(nt, 256, k_pp_edge_keys, tris, nt, keys); }
void emul_corner_keys(const uint32_t *tris, uint32_t nt, unsigned long long *keys) { run(nt, 256, k_pp_corner_keys, tris, nt, keys); }
// flag + (exclusive scan in place of cub) + compaction + row offsets on SORTED keys; returns the number of kept entries
uint32_t emul_csr(const unsigned long long *keys_sorted, uint32_t n, uint32_t nv, uint32_t *row /* nv + 1 */, uint32_t *idx /* n */) {
    uint32_t *flag = (uint32_t *)malloc((size_t)n * 4), *off = (uint32_t *)malloc((size_t)n * 4);
    run(n, 256, k_pp_flag_unique, keys_sorted, n, flag);
    uint32_t acc = 0;
    for (uint32_t k = 0; k < n; ++k) { off[k] = acc; acc += flag[k]; }
    run(n, 256, k_pp_compact_low, keys_sorted, (const uint32_t *)flag, (const uint32_t *)off, n, idx);
    run((uint64_t)nv + 1, 256, k_pp_row_offsets, keys_sorted, (const uint32_t *)off, n, acc, nv, row);
    free(flag); free(off);
    return acc;
}
// the iteration loop of ss_surface_laplacian_smoothing_f32 (pointer swap included); result in verts
void emul_laplacian(uint32_t nv, float *verts, const uint32_t *row, const uint32_t *adj, const float *weights, float beta, uint32_t iterations) {
    float *tmp = (float *)malloc((size_t)nv * 12);
    memcpy(tmp, verts, (size_t)nv * 12);
    float *cur = verts, *buf = tmp;
    for (uint32_t it = 0; it < iterations; ++it) {
        std::swap(cur, buf);
        run(nv, 256, k_pp_laplacian, nv, cur, (const float *)buf, row, adj, weights, beta);
    }
    if (cur != verts) memcpy(verts, cur, (size_t)nv * 12);
    free(tmp);
}
void emul_smooth_normals(uint32_t nv, float *normals, const uint32_t *row, const uint32_t *adj, uint32_t iterations) {
    float *tmp = (float *)malloc((size_t)nv * 12);
    float *a = normals, *b = tmp;
    for (uint32_t it = 0; it < iterations; ++it) { run(nv, 256, k_pp_smooth_normals, nv, (const float *)a, b, row, adj); std::swap(a, b); }
    if (a != normals) memcpy(normals, a, (size_t)nv * 12);
    free(tmp);
}
void emul_area_normals(uint32_t nv, const float *verts, const uint32_t *tris, const uint32_t *row, const uint32_t *inc, float *out) {
    run(nv, 256, k_pp_area_normals, nv, verts, tris, row, inc, out);
}

}  // extern "C"
