"""GPU parity of rvb_obs_hooks / rvb_teacher_record against oracle/hooks_oracle.py.
Bit-exact: dropout decisions, offset, masking, recorder rows, untouched columns.  Noise: <= 1e-5 absolute (logf / cosf / sqrtf
of CUDA's libm against numpy's, a few ulp of a value <= 6 sigma)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import isaac_rover_b200
    return isaac_rover_b200


@pytest.fixture(scope="module")
def HO():
    import hooks_oracle
    return hooks_oracle


def test_deterministic_hooks_bit_exact(R, HO):
    torch.manual_seed(0)
    obs = torch.rand(70, 1750)
    remove_idx = torch.tensor([0, 5, 633, 634, 1745])
    h = R.ObsHooks(1750, offset=0.02, remove_idx=remove_idx)
    got = h.apply(obs.cuda(), epoch=3).cpu().numpy()
    mask = np.zeros(1750, np.uint8)
    mask[(remove_idx + 4).numpy()] = 1
    assert np.array_equal(got, HO.obs_hooks(obs.numpy(), 4, 0.0, 0.0, 0.02, mask, 42, 3))
    ident = R.ObsHooks(1750).apply(obs.cuda(), epoch=1).cpu()
    assert torch.equal(ident, obs)


@pytest.mark.parametrize("N,C,col0", [(1, 1750, 4), (70, 1750, 4), (33, 277, 0), (5, 9, 9)])
def test_noise_dropout_against_oracle(R, HO, N, C, col0):
    torch.manual_seed(N)
    obs = torch.rand(N, C)
    std, p = float(np.float32(0.20 ** 0.5)), 0.1
    h = R.ObsHooks(C, noise_std=std, dropout_p=p, offset=0.02, num_proprioceptive=col0, seed=1234567890123)
    got = h.apply(obs.cuda(), epoch=(5 << 32) + 17, env_offset=1000).cpu().numpy()
    ref = HO.obs_hooks(obs.numpy(), col0, std, p, 0.02, None, 1234567890123, (5 << 32) + 17, env_offset=1000)
    assert np.array_equal(got == np.float32(-0.02), ref == np.float32(-0.02))          # the very same elements dropped
    assert np.abs(got - ref).max() <= 1e-5
    assert np.array_equal(got[:, :col0], ref[:, :col0])


def test_shard_invariance_and_strided_rows(R):
    torch.manual_seed(4)
    wide = torch.rand(64, 1800, device="cuda")
    full = wide.clone()
    h = R.ObsHooks.reference_values(1750)
    h.apply(full[:, :1750], epoch=11)
    assert torch.equal(full[:, 1750:], wide[:, 1750:])                                   # columns beyond C untouched
    a, b = wide[:40].clone(), wide[40:].clone()
    h.apply(a[:, :1750], epoch=11, env_offset=0)
    h.apply(b[:, :1750], epoch=11, env_offset=40)
    assert torch.equal(torch.cat([a, b]), full)
    again = wide.clone()
    h.apply(again[:, :1750], epoch=12)
    assert not torch.equal(again, full)


def test_teacher_recorder_rows_and_file(R, HO, tmp_path):
    torch.manual_seed(2)
    N, C, T = 19, 1750, 3
    rec = R.TeacherRecorder(N, C, 634, 1112, steps=T, directory=str(tmp_path))
    rows = []
    path = None
    for t in range(T):
        obs = torch.rand(N, C, device="cuda")
        actions = torch.rand(N, 2, device="cuda") * 2 - 1
        reset_info = (torch.rand(N, device="cuda") < 0.3).float()
        rows.append(HO.teacher_row(reset_info.cpu().numpy(), actions.cpu().numpy(), obs.cpu().numpy()))
        path = rec.record(reset_info, actions, obs)
        assert (path is None) == (t < T - 1)
    saved = torch.load(path)
    assert saved["info"] == {"reset": 1, "actions": 2, "proprioceptive": 4, "sparse": 634, "dense": 1112}      # rover.py:304-310
    assert saved["data"].shape == (T, N, 3 + C)
    assert np.array_equal(saved["data"].numpy(), np.stack(rows))
    assert rec.curr_timestep == 0 and rec.dataset_nr == 1 and os.path.basename(path) == "teacher_dataset_0.pt"


def test_rovertask_applies_hooks_and_records(R):
    """get_observations with the switches on: the recorder sees the PREVIOUS obs_buf (rover.py:299 runs before :320-325) and
    the hooks run last (rover.py:326-329)."""
    dev = "cuda:0"
    w = R.synth.make_world(length=12.0, nv=44, K=64, n_stones=8, seed=3, build_index=None)
    w.map_indices = R.build_knn_index(w.triangles, w.vertices, w.G, w.res, w.K, device=dev).cpu()
    kr = min(w.K, w.rock_triangles.shape[0])
    w.rock_indices = R.build_knn_index(w.rock_triangles, w.rock_vertices, w.G, w.res, kr, device=dev).cpu()
    st = R.synth.make_env_state(w, 16, seed=5, margin=3.0)
    task = R.synth.make_task(w, st, device=dev, level=2)
    task.get_observations()
    plain = task.obs_buf.clone()
    task.obs_hooks = R.ObsHooks(1750, offset=0.02, remove_idx=[0, 1, 2])
    task.save_teacher_data = True
    task.teacher_recorder = R.TeacherRecorder(16, 1750, 634, 1112, steps=4, save=False)
    task.reset_info[:] = 1
    task._teacher_actions = st["actions"].to(dev)[:, :2].float()
    before = task.obs_buf.clone()
    task.get_observations()
    row = task.teacher_recorder.teacher_dataset[0]
    assert torch.equal(row[:, 3:], before) and torch.equal(row[:, 0], task.reset_info) and torch.equal(row[:, 1:3], task._teacher_actions)
    expect = plain - 0.02
    expect[:, 4:7] = 0
    assert torch.equal(task.obs_buf, expect)


def test_fused_and_unfused_hot_step_with_hooks_agree_over_steps(R):
    """hot_step(fused=True) and hot_step(fused=False) with noise / dropout hooks and the teacher recorder switched on: the same
    observations, the same recorder rows, a NEW Philox epoch every step (global_step advances on both paths), and the device-side
    reset path sets reset_info like rover.py:420-422."""
    dev = "cuda:0"
    w = R.synth.make_world(length=12.0, nv=44, K=64, n_stones=8, seed=3, build_index=None)
    w.map_indices = R.build_knn_index(w.triangles, w.vertices, w.G, w.res, w.K, device=dev).cpu()
    kr = min(w.K, w.rock_triangles.shape[0])
    w.rock_indices = R.build_knn_index(w.rock_triangles, w.rock_vertices, w.G, w.res, kr, device=dev).cpu()
    st = R.synth.make_env_state(w, 24, seed=5, margin=3.0)
    runs = []
    for fused in (True, False):
        task = R.synth.make_task(w, st, device=dev, level=2)
        task.obs_hooks = R.ObsHooks(1750, noise_std=0.1, dropout_p=0.05, offset=0.02, remove_idx=[3, 9])
        task.save_teacher_data = True
        task.teacher_recorder = R.TeacherRecorder(24, 1750, 634, 1112, steps=8, save=False)
        obs_steps = []
        for k in range(3):
            act = (st["actions"] * (1.0 - 0.25 * k)).to(dev)
            obs, rew, reset, _ = task.hot_step(act, fused=fused, device_reset=(k == 2))
            obs_steps.append((obs.clone(), rew.clone(), reset.clone()))
        assert task.global_step == 3
        runs.append((obs_steps, task.teacher_recorder.teacher_dataset[:3].clone(), task.reset_info.clone()))
    (fa, ra, ia), (fb, rb, ib) = runs
    for (oa, wa, sa), (ob, wb, sb) in zip(fa, fb):
        assert torch.equal(oa, ob) and torch.equal(sa, sb)
        assert (wa - wb).abs().max().item() <= 1e-6
    assert torch.equal(ra, rb) and torch.equal(ia, ib)
    assert bool((ia == 1).all())                                   # some env resets in the synthetic state -> every env flagged
    # the noise pattern changes from step to step (a frozen epoch would repeat it): heightmap columns of step 0 and step 1 come
    # from the same poses, so their difference is pure noise / dropout
    d01 = (fa[0][0][:, 4:] - fa[1][0][:, 4:]).abs()
    assert float((d01 > 0).float().mean()) > 0.5
