"""GPU parity for subdomain sizes whose point count is not 8k+1 (partial last brick, no extension bricks), against the pinned
oracle, and the frame-sequence API."""
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


@pytest.mark.gpu
def test_cuda_frame_stream_matches_serial_calls(ss):
    """distributed.FrameStream (two frames in flight: upload of frame i+1 and download of frame i-1 during the compute of frame i)
    returns, frame by frame, exactly what serial reconstruct_surface calls return -- different clouds per frame, so a mixed-up
    buffer would show."""
    import torch
    from splashsurf_b200 import distributed as ssd, synthetic as syn
    kw = dict(particle_radius=0.025, smoothing_length=2.0, cube_size=0.5)
    frames = [syn.jittered_cube(14 + 2 * k, 0.025, 900 + k) for k in range(5)]
    ctx = ss.Context()
    try:
        want = [ss.reconstruct_surface(p, context=ctx, **kw) for p in frames]
        fs = ssd.FrameStream(ctx, ss.make_params(**kw))
        pinned = [torch.from_numpy(p).pin_memory() for p in frames]
        got = []
        fs.submit(pinned[0])
        for k in range(len(frames)):
            if k + 1 < len(frames):
                fs.submit(pinned[k + 1])
            prev = fs.advance()
            if prev is not None:
                got.append((prev["vertices"].numpy().copy(), prev["triangles"].numpy().copy()))
        last = fs.drain()
        got.append((last["vertices"].numpy().copy(), last["triangles"].numpy().copy()))
        fs.close()
    finally:
        ctx.close()
    assert len(got) == len(frames)
    for (v, t), w in zip(got, want):
        assert np.array_equal(v, w.mesh.vertices) and np.array_equal(t.astype(np.uint64), w.mesh.triangles)
