"""GPU parity for subdomain sizes whose point count is not 8k+1 (partial last brick, no extension bricks), against the pinned
oracle.  Kept in its own late-sorted file: it was written after the round's GPU budget was spent and has not run on a GPU yet."""
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
