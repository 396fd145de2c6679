"""CPU: small pieces of host logic that need no GPU."""

import numpy as np

from pyslam_b200 import remap_instance_ids
from pyslam_b200 import sharding


def test_remap_instance_ids_matches_a_plain_loop():
    """volumetric.remap_instance_ids (cpp/volumetric/image_utils.h:69-163): mapped ids are replaced, everything
    else - and everything when the map is empty - becomes the invalid id."""
    rng = np.random.default_rng(0)
    img = rng.integers(-1, 40, size=(37, 53)).astype(np.int32)
    m = {int(k): int(v) for k, v in zip(rng.choice(40, 15, replace=False), rng.integers(-1, 500, 15))}
    m[0] = 0
    out = remap_instance_ids(img, m)
    ref = np.array([[m.get(int(v), -1) for v in row] for row in img], np.int32)
    assert out.dtype == np.int32 and np.array_equal(out, ref)
    assert np.all(remap_instance_ids(img, {}) == -1)
    assert np.all(remap_instance_ids(img, {}, invalid_instance_id=-7) == -7)
    assert remap_instance_ids(np.zeros((0, 0), np.int32), m).size == 0


def test_block_key_hash_matches_the_reference_formula():
    """BlockKeyHash = h1 ^ (h2 << 1) ^ (h3 << 2) on sign-extended 64-bit values (voxel_hashing.h:106-113);
    SURVEY.md 8c known answer: block (-1, 12, 30) -> 2^64 - 97."""
    assert int(sharding.block_key_hash(np.array([[-1, 12, 30]]))[0]) == 2 ** 64 - 97
    keys = np.array([[0, 0, 0], [1, 2, 3], [-5, 7, -9], [2 ** 20, -2 ** 20, 5]], np.int32)
    exp = [(int(x) & (2 ** 64 - 1)) ^ ((int(y) << 1) & (2 ** 64 - 1)) ^ ((int(z) << 2) & (2 ** 64 - 1))
           for x, y, z in keys.astype(np.int64)]
    assert [int(h) for h in sharding.block_key_hash(keys)] == exp
    owners = sharding.owner_of(keys, 8)
    assert owners.tolist() == [e % 8 for e in exp]
