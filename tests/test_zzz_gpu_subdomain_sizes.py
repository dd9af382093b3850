"""GPU parity for subdomain sizes whose point count is not 8k+1 (partial last brick, no extension bricks), against the pinned
oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _parity(oracle_mod, g, o, S=64):
    return oracle_mod.mesh_parity(g.mesh.vertices, g.mesh.triangles, g.vertex_edge_keys, o["vertices"], o["triangles"], o["vertex_keys"], S)


@pytest.mark.parametrize("S", [20, 50])
def test_cuda_subdomain_size_not_multiple_of_8(ss, oracle_mod, S):
    """Tiles whose point count is not 8k+1 take the level-set path without extension bricks (partial last brick)."""
    from splashsurf_b200 import synthetic as syn
    p = syn.splash((18, 18, 18), 4, 0.025, 140 + S)
    kw = dict(particle_radius=0.025, smoothing_length=2.0, cube_size=0.6, subdomain_num_cubes_per_dim=S, subdomain_grid_auto_disable=False)
    o = oracle_mod.reconstruct(p, **kw)
    g = ss.reconstruct_surface(p, with_debug=True, **kw)
    assert np.array_equal(g.particle_densities, o["particle_densities"])
    m = _parity(oracle_mod, g, o, S)
    assert m["keys_equal"] and m["triangles_equal"] and m["n_not_bitexact"] == 0, m



def test_cuda_reusable_host_buffers(ss, oracle_mod):
    """Context.reuse_host_buffers: the result arrays are views of the context's page-locked buffers -- same values as the default
    (freshly allocated) arrays, and overwritten by the next reconstruction on the context."""
    from splashsurf_b200 import synthetic as syn
    kw = dict(particle_radius=0.025, smoothing_length=2.0, cube_size=0.5)
    a, b = syn.jittered_cube(14, 0.025, 31), syn.jittered_cube(12, 0.025, 32)
    ctx = ss.Context()
    try:
        want_a, want_b = ss.reconstruct_surface(a, context=ctx, **kw), ss.reconstruct_surface(b, context=ctx, **kw)
        ctx.reuse_host_buffers = True
        ra = ss.reconstruct_surface(a, context=ctx, **kw)
        assert np.array_equal(ra.mesh.vertices, want_a.mesh.vertices) and np.array_equal(ra.mesh.triangles, want_a.mesh.triangles)
        assert np.array_equal(ra.particle_densities, want_a.particle_densities)
        keep = ra.mesh.vertices.copy()
        rb = ss.reconstruct_surface(b, context=ctx, **kw)
        assert np.array_equal(rb.mesh.vertices, want_b.mesh.vertices) and np.array_equal(rb.mesh.triangles, want_b.mesh.triangles)
        assert np.array_equal(keep, want_a.mesh.vertices)
    finally:
        ctx.close()
