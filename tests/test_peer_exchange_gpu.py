"""GPU (one device is enough): the in-kernel exchange of the advantage moments (ppo_gae_normalized_peers) with the peer
played by the test.  Rank 0 of a world of two runs on cuda:0; "rank 1" is a second mailbox on the same device whose
record and flag the test either delivers (the launch must normalise with the moments of both ranks, alternating the
two halves of the mailbox by epoch parity) or withholds (the launch must give up after the time-out, store the epoch
into the status word and leave NaN advantages -- never hang, never return garbage silently)."""
import ctypes

import numpy as np
import pytest
import torch

import oracle
from ppo_b200 import _lib, rollout, synthetic

pytestmark = pytest.mark.gpu

STEPS, ENVS = 64, 96
GAMMA, LAMBDA, EPS = 0.999, 0.95, 1e-8


class _TwoRanks:
    """Rank 0's launch arguments with both mailboxes on this device."""

    def __init__(self):
        self.lib = _lib.load()
        dev = torch.device("cuda", 0)
        r, d, v = synthetic.rollout(STEPS, ENVS, seed=3)
        self.np_in = (r, d, v)
        self.r = torch.from_numpy(r).to(dev)
        self.d = torch.from_numpy(d.astype(np.uint8)).to(dev)
        self.v = torch.from_numpy(v).to(dev)
        words = 2 * 2 * _lib.STATS_LEN + 2 * 2  # [2][world][16] doubles + [2][world] flags
        self.box = [torch.zeros(words, dtype=torch.float64, device=dev) for _ in range(2)]
        mail_bytes = 2 * 2 * _lib.STATS_LEN * 8
        self.mail = (ctypes.c_void_p * 2)(*[b.data_ptr() for b in self.box])
        self.flags = (ctypes.c_void_p * 2)(*[b.data_ptr() + mail_bytes for b in self.box])
        self.status = torch.zeros(1, dtype=torch.int64, device=dev)
        self.adv = torch.empty((STEPS, ENVS), dtype=torch.float32, device=dev)
        self.ret = torch.empty_like(self.adv)
        self.norm = torch.empty(STEPS * ENVS, dtype=torch.float32, device=dev)
        self.stats = torch.zeros(_lib.STATS_LEN, dtype=torch.float64, device=dev)
        self.gstats = torch.zeros(_lib.STATS_LEN, dtype=torch.float64, device=dev)
        self.ws = torch.zeros(max(int(self.lib.ppo_gae_workspace_bytes(ENVS)), 4096), dtype=torch.uint8, device=dev)
        assert self.lib.ppo_gae_normalized_fits(STEPS, ENVS, _lib.DONE_U8, 1) == 1

    def deliver(self, epoch, record):
        """What rank 1's folding CTA does: its record into slot 1 of rank 0's mailbox (the epoch's half), then the flag."""
        half = (epoch & 1) * 2
        own = self.box[0]
        own[(half + 1) * _lib.STATS_LEN:(half + 2) * _lib.STATS_LEN] = record
        flags = own[2 * 2 * _lib.STATS_LEN:].view(torch.int64)
        flags[half + 1] = epoch

    def launch(self, epoch, timeout_ms):
        r, d, v = self.r, self.d, self.v
        rc = self.lib.ppo_gae_normalized_peers(
            r.data_ptr(), r.stride(0), v.data_ptr(), v.stride(0), v.stride(1), v[STEPS].data_ptr(), v.stride(1),
            d.data_ptr(), _lib.DONE_U8, d.stride(0), STEPS, ENVS, GAMMA, LAMBDA, LAMBDA, EPS, 0.0,
            self.adv.data_ptr(), self.ret.data_ptr(), self.norm.data_ptr(), self.stats.data_ptr(),
            self.gstats.data_ptr(), self.ws.data_ptr(), self.ws.numel(), 0, 2, self.mail, self.flags, epoch,
            self.status.data_ptr(), float(timeout_ms), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "ppo_gae_normalized_peers")
        assert rc == 1
        torch.cuda.synchronize()


def test_withheld_peer_times_out_with_a_status_word():
    t = _TwoRanks()
    t.launch(epoch=1, timeout_ms=50.0)  # rank 1 never raises its flag
    assert int(t.status.item()) == 1
    assert torch.isnan(t.norm).all()
    # the raw advantages and the local record do not depend on the exchange
    r, d, v = t.np_in
    adv = oracle.gae(r, v[:-1], v[-1], d, GAMMA, LAMBDA)
    assert np.array_equal(t.adv.cpu().numpy(), adv)
    assert t.stats[_lib.STAT_COUNT].item() == STEPS * ENVS


def test_delivered_peer_records_alternate_halves():
    t = _TwoRanks()
    r, d, v = t.np_in
    adv = oracle.gae(r, v[:-1], v[-1], d, GAMMA, LAMBDA).astype(np.float64)
    for epoch in (1, 2, 3):
        # rank 1's shard: a record with a different mean / spread every epoch
        other = np.random.RandomState(epoch).normal(0.3 * epoch, 1.0 + epoch, size=1000)
        rec = torch.zeros(_lib.STATS_LEN, dtype=torch.float64)
        rec[_lib.STAT_COUNT], rec[_lib.STAT_ADV_SUM], rec[_lib.STAT_ADV_SUMSQ] = other.size, other.sum(), (other * other).sum()
        rec[_lib.STAT_ADV_MIN], rec[_lib.STAT_ADV_MAX] = other.min(), other.max()
        rec[_lib.STAT_RET_MIN], rec[_lib.STAT_RET_MAX] = np.inf, -np.inf
        t.deliver(epoch, rec.cuda())
        torch.cuda.synchronize()
        t.launch(epoch, timeout_ms=2000.0)
        assert int(t.status.item()) == 0
        n = adv.size + other.size
        mean = (adv.sum() + other.sum()) / n
        var = max(((adv * adv).sum() + (other * other).sum()) / n - mean * mean, 0.0)
        want = ((adv.astype(np.float32) - np.float32(mean)) / (np.float32(np.sqrt(var)) + np.float32(EPS))).reshape(-1)
        got = t.norm.cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-6)
        assert t.gstats[_lib.STAT_COUNT].item() == n
        # the own record went into the epoch's half of BOTH mailboxes, the other half keeps the previous epoch
        half = (epoch & 1) * 2
        for box in t.box:
            assert box[half * _lib.STATS_LEN + _lib.STAT_COUNT].item() == STEPS * ENVS
            assert box[2 * 2 * _lib.STATS_LEN:].view(torch.int64)[half].item() == epoch


def test_return_pass_reports_a_timed_out_exchange():
    """ReturnPass surfaces the status word as TimeoutError (host side only: a mailbox with a set status word)."""
    box = rollout.PeerMailbox.__new__(rollout.PeerMailbox)
    box.status = torch.zeros(1, dtype=torch.int64)
    box.rank, box.epoch, box.timeout_ms = 0, 7, 4000.0
    box.check()
    box.status[0] = 7
    with pytest.raises(TimeoutError, match="pass 7"):
        box.check()
