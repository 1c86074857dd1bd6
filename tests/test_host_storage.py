"""CPU: host logic of RolloutStorage (algorithms/algo_utils/storage.py:7-138 in the reference) that needs no kernel — buffer
layout, in-place observation slots, overflow error, sampler geometry (SURVEY KAT-3), DAgger ring cursors."""
import numpy as np
import os
import pytest
import torch

from partmanip_b200.algorithms.algo_utils import storage as S

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampler_geometry.npz"))


def test_ppo_buffers_layout_and_dtypes():
    st = S.RolloutStorage(4, 3, 6, 2, "cpu", default_succ_value=500, whole_adv_norm=True)
    shapes = {"observations": (3, 4, 6), "actions": (3, 4, 2), "mu": (3, 4, 2), "sigma": (3, 4, 2)}
    for name in ("rewards", "dones", "succs", "actions_log_prob", "values", "returns", "advantages", "step_id"):
        shapes[name] = (3, 4, 1)
    for name, shape in shapes.items():
        t = getattr(st, name)
        assert tuple(t.shape) == shape and t.dtype == (torch.bool if name in ("dones", "succs") else torch.float32), name
        assert not t.any()
    assert st.cur_buf_size == 12 and st.step == 0 and st.default_succ_value == 500 and st.whole_adv_norm


def test_add_transitions_in_place_slot_and_overflow():
    st = S.RolloutStorage(4, 2, 6, 2, "cpu")
    g = torch.Generator().manual_seed(0)
    for t in range(2):
        st.obs_slot().copy_(torch.full((4, 6), float(t + 1)))                 # the producer writes the slot directly: no copy kernel
        st.add_transitions(st.obs_slot(), torch.rand(4, 2, generator=g), torch.full((4,), 0.5), torch.tensor([0, 1, 0, 0]).bool(),
                           torch.zeros(4, 1).bool(), torch.full((4, 1), 2.0), torch.full((4,), -1.0), torch.zeros(4, 2), torch.ones(4, 2))
    assert st.step == 2 and float(st.observations[1].mean()) == 2.0
    assert st.rewards.shape == (2, 4, 1) and float(st.rewards.sum()) == 4.0 and int(st.dones.sum()) == 2
    assert float(st.actions_log_prob.sum()) == -8.0 and float(st.sigma.sum()) == 16.0
    with pytest.raises(AssertionError, match="Rollout buffer overflow"):
        st.obs_slot()
    with pytest.raises(AssertionError, match="Rollout buffer overflow"):
        st.add_transitions(*[torch.zeros(4, 6)] * 9)
    st.clear()
    assert st.step == 0


def test_sampler_geometry_matches_reference_recording():
    """rows of the recording: (E, T, n_minibatches, number of batches, batch size, first index, last index) of the reference's
    BatchSampler(SequentialSampler) — ours yields the same batches as contiguous ranges."""
    for E, T, nmb, count, size, first, last in G["geo"].tolist():
        st = S.RolloutStorage.__new__(S.RolloutStorage)
        st.cur_buf_size, st.sampler, st.device = E * T, "sequential", "cpu"
        batches = st.mini_batch_generator(nmb)
        assert len(batches) == count and all(len(b) == size for b in batches)
        assert [b.start for b in batches] == [k * size for k in range(count)]           # contiguous slices of the flat buffer
        assert batches[0][0] == first and batches[-1][len(batches[-1]) - 1] == last


def test_sampler_geometry_kat3():
    st = S.RolloutStorage.__new__(S.RolloutStorage)
    st.sampler, st.device = "sequential", "cpu"
    for (E, T, nmb), (count, size) in {(64, 8, 8): (8, 64), (2048, 8, 8): (8, 2048), (4096, 8, 8): (16, 2048)}.items():
        st.cur_buf_size = E * T
        b = st.mini_batch_generator(nmb)
        assert (len(b), len(b[0])) == (count, size)
    st.sampler = "random"
    st.cur_buf_size = 100
    rb = st.mini_batch_generator(3)
    seen = torch.cat(list(rb))
    assert len(rb) == 3 and seen.numel() == 99 and seen.unique().numel() == 99            # a permutation, drop_last
    st.sampler = "bogus"
    with pytest.raises(NotImplementedError):
        st.mini_batch_generator(3)


def test_dagger_ring_cursors(monkeypatch):
    monkeypatch.setattr(S.ops, "copy_rows", lambda src, dst: dst.copy_(src))
    st = S.RolloutStorage(4, 3, 6, 2, "cpu", tea_obs_shape=5, max_length=7)
    assert st.observations.shape == (12, 6) and st.tea_obs.shape == (12, 5) and st.succ_buf_ind == 28 and st.cur_buf_size == 0
    for i in range(4):
        st.add_transitions_dagger(torch.full((4, 6), float(i)), torch.full((4, 5), float(10 + i)))
    assert st.cur_buf_size == 12 and st.mix_buf_ind == 4                               # wrapped: step 3 overwrote rows 0..3
    assert float(st.observations[0, 0]) == 3.0 and float(st.observations[4, 0]) == 1.0 and float(st.tea_obs[0, 0]) == 13.0


def test_random_sampler_reproduces_the_reference_permutation():
    """storage.py:133-137: BatchSampler(SubsetRandomSampler(range(n)), size, drop_last=True) draws torch.randperm(n) from the host
    default generator; for the same torch.manual_seed ours yields the same index batches (two epochs = two fresh permutations)."""
    from torch.utils.data.sampler import BatchSampler, SubsetRandomSampler
    st = S.RolloutStorage.__new__(S.RolloutStorage)
    st.sampler, st.device, st.cur_buf_size = "random", "cpu", 8 * 37
    torch.manual_seed(123)
    ref = BatchSampler(SubsetRandomSampler(range(st.cur_buf_size)), min(st.cur_buf_size // 8, 2048), drop_last=True)
    want = [list(b) for _ in range(2) for b in ref]
    torch.manual_seed(123)
    ours = st.mini_batch_generator(8)
    got = [b.tolist() for _ in range(2) for b in ours]
    assert got == want
