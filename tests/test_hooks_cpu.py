"""Observation hooks / teacher recorder: the oracle against the reference's own statements
(rover.py:299,326-329,364,374-375 run literally on torch tensors) and the distributional properties of its Philox draws."""
import numpy as np
import pytest
import torch

import hooks_oracle as HO


def test_deterministic_hooks_equal_the_reference_statements():
    torch.manual_seed(0)
    obs_buf = torch.rand(37, 1750)
    remove_idx = torch.tensor([0, 5, 633, 634, 1745])
    ref = obs_buf.clone()
    ref = ref - 0.02                                   # rover.py:328
    ref[:, remove_idx + 4] = 0                         # rover.py:329
    mask = np.zeros(1750, np.uint8)
    mask[(remove_idx + 4).numpy()] = 1
    got = HO.obs_hooks(obs_buf.numpy(), 4, 0.0, 0.0, 0.02, mask, seed=1, epoch=2)
    assert np.array_equal(got, ref.numpy())
    # all switches off = identity
    assert np.array_equal(HO.obs_hooks(obs_buf.numpy(), 4, 0.0, 0.0, 0.0, None, 1, 2), obs_buf.numpy())


def test_teacher_row_equals_the_reference_statements():
    torch.manual_seed(1)
    N, C = 11, 1750
    obs_buf, actions, reset_info = torch.rand(N, C), torch.rand(N, 2) * 2 - 1, (torch.rand(N) < 0.3).float()
    data_curr_timestep = torch.empty((N, 2 + C + 1))   # rover.py:176
    data_curr_timestep[:, 0] = reset_info[:]           # rover.py:364
    data_curr_timestep[:, 1] = actions[:, 0]           # rover.py:374
    data_curr_timestep[:, 2] = actions[:, 1]           # rover.py:375
    data_curr_timestep[:, 3:] = obs_buf                # rover.py:299
    assert np.array_equal(HO.teacher_row(reset_info.numpy(), actions.numpy(), obs_buf.numpy()), data_curr_timestep.numpy())


def test_noise_and_dropout_distributions():
    N, C = 256, 1750
    base = np.full((N, C), 0.25, np.float32)
    std = float(np.float32(0.20 ** 0.5))
    out = HO.obs_hooks(base, 4, std, 0.0, 0.0, None, seed=42, epoch=7)
    assert np.array_equal(out[:, :4], base[:, :4])                      # proprioceptive columns untouched (rover.py:326)
    z = (out[:, 4:] - 0.25) / std
    n = z.size
    assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.std() - 1) < 0.01
    assert abs((np.abs(z) < 1).mean() - 0.6827) < 0.005 and np.abs(z).max() < 6.0
    assert abs(np.corrcoef(z[:, :-1].ravel(), z[:, 1:].ravel())[0, 1]) < 0.01      # neighbouring columns independent
    out = HO.obs_hooks(base, 4, 0.0, 0.1, 0.0, None, seed=42, epoch=7)
    dropped = out[:, 4:] == 0
    assert abs(dropped.mean() - 0.1) < 0.003
    assert np.allclose(out[:, 4:][~dropped], 0.25 / 0.9, rtol=1e-6)    # F.dropout scaling of the survivors
    # the expected value is preserved, like F.dropout
    assert abs(out[:, 4:].mean() - 0.25) < 0.002


def test_draws_depend_on_seed_epoch_env_column_only():
    a = HO.obs_hooks(np.zeros((8, 40), np.float32), 4, 1.0, 0.0, 0.0, None, seed=3, epoch=9, env_offset=0)
    b = HO.obs_hooks(np.zeros((4, 40), np.float32), 4, 1.0, 0.0, 0.0, None, seed=3, epoch=9, env_offset=4)
    assert np.array_equal(a[4:], b)                                     # shard = slice of the unsharded run
    c = HO.obs_hooks(np.zeros((8, 40), np.float32), 4, 1.0, 0.0, 0.0, None, seed=3, epoch=10)
    d = HO.obs_hooks(np.zeros((8, 40), np.float32), 4, 1.0, 0.0, 0.0, None, seed=4, epoch=9)
    assert not np.array_equal(a, c) and not np.array_equal(a, d)
    assert len(np.unique(a[:, 4:])) == a[:, 4:].size


def test_hooks_have_no_cpu_path():
    import isaac_rover_b200 as R
    with pytest.raises(RuntimeError):
        R.ObsHooks(1750, device="cpu")
    with pytest.raises(RuntimeError):
        R.TeacherRecorder(4, 1750, 634, 1112, device="cpu")
