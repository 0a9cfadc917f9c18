"""update_aabb (SURVEY 8f "next #2"): oracle restatement vs the double-precision host fixture on CPU; CUDA kernel vs oracle on GPU."""
import numpy as np
import pytest

from avian_b200 import api, scenes
from avian_b200.fixture import HostPipeline

import oracle_lib


def _colliders(n=3000, seed=5, scalar=np.float32):
    rng = np.random.default_rng(seed)
    s = np.dtype(scalar)
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    c = api.Colliders(shape=rng.choice([0, 0, 1], size=n).astype(np.uint8), dims=rng.uniform(0.1, 2.0, size=(n, 3)).astype(s),
                      position=(rng.normal(size=(n, 3)) * 20).astype(s), rotation=q.astype(s),
                      linear_velocity=(rng.normal(size=(n, 3)) * 5).astype(s), angular_velocity=(rng.normal(size=(n, 3)) * 3).astype(s),
                      collision_margin=rng.choice([0.0, 0.0, 0.05], size=n).astype(s),
                      speculative_margin=rng.choice([np.inf, np.inf, 0.0, 0.2], size=n).astype(s))
    c.linear_velocity[::11] = 0
    c.angular_velocity[::7] = 0
    return c


def test_oracle_matches_host_fixture_on_scene():
    """default configuration (speculative margin MAX, no collision margin): the oracle's f32 restatement agrees with the fixture's
    double-precision AABBs to f32 accuracy, and every swept AABB contains the box at its start pose"""
    sc = scenes.cube_stack(5, 4, 5, brick=True)
    rng = np.random.default_rng(1)
    sc.bodies.linear_velocity[1:] = rng.normal(size=(sc.bodies.count - 1, 3)).astype(np.float32)
    sc.bodies.angular_velocity[1:] = rng.normal(size=(sc.bodies.count - 1, 3)).astype(np.float32)
    pipe = HostPipeline(sc.shape_type, sc.dims, sc.friction, sc.restitution)
    dt = 1.0 / 60.0
    mn_f, mx_f = pipe.update_aabbs(sc.bodies, dt)
    c = api.Colliders(shape=sc.shape_type.astype(np.uint8), dims=sc.dims.astype(np.float32), position=sc.bodies.position, rotation=sc.bodies.rotation,
                      linear_velocity=sc.bodies.linear_velocity, angular_velocity=sc.bodies.angular_velocity)
    oracle_lib.update_aabbs(api.aabb_params(dt), c)
    assert np.abs(c.aabb_min - mn_f).max() < 2e-5 and np.abs(c.aabb_max - mx_f).max() < 2e-5
    assert (c.aabb_min <= sc.bodies.position - 0.5).all() and (c.aabb_max >= sc.bodies.position[:, :] + 0.5)[1:].all()


@pytest.mark.gpu
@pytest.mark.parametrize("scalar", [np.float32, np.float64])
def test_gpu_update_aabbs_bit_exact(scalar):
    with api.Context(device=0, scalar=scalar) as ctx:
        c, co = _colliders(scalar=scalar), _colliders(scalar=scalar)
        prm = api.aabb_params(1.0 / 60.0)
        oracle_lib.update_aabbs(prm, co)
        ctx.update_aabbs(prm, c)
        assert np.isfinite(c.aabb_min).all() and (c.aabb_min <= c.aabb_max).all()
        assert np.array_equal(c.aabb_min, co.aabb_min) and np.array_equal(c.aabb_max, co.aabb_max)
        # finite default speculative margin
        prm2 = api.aabb_params(1.0 / 60.0, default_speculative_margin=0.1)
        c.speculative_margin = None; co.speculative_margin = None
        oracle_lib.update_aabbs(prm2, co)
        ctx.update_aabbs(prm2, c)
        assert np.array_equal(c.aabb_min, co.aabb_min) and np.array_equal(c.aabb_max, co.aabb_max)
