"""Host-side schedule logic of the fast loop against the reference's control flow (no GPU):

  * ``epoch_schedule``: optimiser-step counts of both training phases equal the counts of the reference's nested
    ``while`` / ``for`` loop (training/training.py:87-178), including the counts recorded from real reference runs
    (tests/golden/psnr_configs.json, ``optimiser_steps``) and the advisor's counter-examples
    (max_pass = 12 -> 8 + 4 passes, max_pass = 3 -> phase 1 runs 2 passes);
  * ``LRSchedule``: the same decisions as NeurcompDecayStrategy / SmallifyDecayStrategy
    (training/learning_rate_decay.py:21-57) on a recorded loss sequence, across both phases.
"""
import json
import math
import os

import pytest

from latent_feature_grid_compression_b200.training.fast_loop import LRSchedule, epoch_schedule

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def reference_steps(n_voxels, batch_size, sample_size, max_pass):
    """Structural transcription of solve_model's loop skeleton (training/training.py:76-114,178)."""
    voxel_seen, volume_passes, steps = 0.0, 0.0, 0
    loader_len = math.ceil(n_voxels / batch_size)               # len(DataLoader(dataset, batch_size)), dataset len n_voxels
    while int(volume_passes) + 1 < max_pass:
        for idx in range(loader_len):
            steps += 1
            # the last batch of an epoch is partial (drop_last=False) and :108 counts the samples actually seen;
            # pinned by the 60-pass reference run of tests/golden/psnr_configs.json (30365 steps, not 30362)
            voxel_seen += min(batch_size, n_voxels - idx * batch_size) * sample_size
            volume_passes = voxel_seen / n_voxels
            if int(volume_passes) >= max_pass:
                break
    return steps, volume_passes


@pytest.mark.parametrize('n_voxels,batch_size,sample_size,max_pass', [
    (150 ** 3, 1024, 16, 50), (255 ** 3, 2048, 16, 60), (48 ** 3, 256, 16, 12), (48 ** 3, 256, 16, 3),
    (48 ** 3, 256, 2, 40), (31 * 29 * 37, 64, 16, 7), (1024 ** 3, 16384, 16, 2)])
def test_epoch_schedule_counts_match_the_reference_loop(n_voxels, batch_size, sample_size, max_pass):
    for frac in (2.0 / 3.0, 1.0 / 3.0):
        mp = max_pass * frac
        got = list(epoch_schedule(n_voxels, batch_size, sample_size, mp))
        want, want_passes = reference_steps(n_voxels, batch_size, sample_size, mp)
        assert len(got) == want
        if got:
            assert got[-1][2] == want_passes
            assert [g[0] for g in got] == list(range(1, want + 1))
            # prior / current pass numbers are what the decay strategies are fed with
            assert all(got[i][1] == int(got[i - 1][2]) for i in range(1, len(got))) and got[0][1] == 0


def test_advisor_counter_examples():
    """with sample_size = 16 one DataLoader epoch is 16 passes, so the reference reaches max_pass."""
    nv = 48 ** 3
    p1 = list(epoch_schedule(nv, 256, 16, 12 * 2.0 / 3.0))
    p2 = list(epoch_schedule(nv, 256, 16, 12 * 1.0 / 3.0))
    assert int(p1[-1][2]) == 8 and int(p2[-1][2]) == 4
    p1 = list(epoch_schedule(nv, 256, 16, 3 * 2.0 / 3.0))
    assert int(p1[-1][2]) == 2
    assert list(epoch_schedule(nv, 256, 16, 1.0)) == []          # int(0) + 1 < 1.0 is false: no step at all


def test_step_counts_of_real_reference_runs():
    path = os.path.join(GOLD, 'psnr_configs.json')
    if not os.path.exists(path):
        pytest.skip('no config goldens')
    recs = [r for r in json.load(open(path)) if 'optimiser_steps' in r]
    if not recs:
        pytest.skip('goldens carry no step counts')
    for r in recs:
        a = r['args']
        R = int(r['volume'].split('(')[1].rstrip(')'))
        n = sum(len(list(epoch_schedule(R ** 3, a['batch_size'], a['sample_size'], r['max_pass'] * f)))
                for f in (2.0 / 3.0, 1.0 / 3.0))
        assert n == r['optimiser_steps'], (r['config'], r['seed'], n, r['optimiser_steps'])


class _Opt:
    def __init__(self, lr):
        self.param_groups = [{'lr': lr}]


class _RefNeurcomp:        # training/learning_rate_decay.py:21-33
    def __init__(self, opt, pass_decay, lr_decay):
        self.optimizer, self.epoch_delay, self.lr_decay = opt, pass_decay, lr_decay

    def decay_learning_rate(self, prior, cur, loss=0):
        if prior != int(cur) and (int(cur) + 1) % self.epoch_delay == 0:
            for g in self.optimizer.param_groups:
                g['lr'] *= self.lr_decay
        return False


class _RefPlateau:         # training/learning_rate_decay.py:36-57
    def __init__(self, opt, smallify_decay, lr_decay, lr_stop=1e-07):
        self.optimizer, self.epoch_delay, self.lr_decay, self.lr_stop = opt, smallify_decay, lr_decay, lr_stop
        self.last_loss, self.no_gain_epoch = None, 0

    def decay_learning_rate(self, prior, cur, loss=0):
        if prior != int(cur):
            if self.last_loss is None or loss < self.last_loss:
                self.last_loss, self.no_gain_epoch = loss, 0
            else:
                self.no_gain_epoch += 1
            if self.no_gain_epoch == self.epoch_delay:
                for g in self.optimizer.param_groups:
                    if g['lr'] > self.lr_stop:
                        g['lr'] *= self.lr_decay
                    else:
                        return True
                self.no_gain_epoch = 0
            return False


class _Trainer:
    def __init__(self):
        self.lr = None

    def set_lr(self, lr):
        self.lr = lr


@pytest.mark.parametrize('smallify_decay', [0, 2])
def test_lr_schedule_matches_the_reference_strategies_across_both_phases(smallify_decay):
    import random
    rng = random.Random(5)
    args = dict(pass_decay=3, lr_decay=0.1, smallify_decay=smallify_decay)
    nv, bs, ss = 31 * 29 * 37, 64, 16
    opt = _Opt(0.008)
    ref = _RefPlateau(opt, smallify_decay, 0.1) if smallify_decay else _RefNeurcomp(opt, 3, 0.1)
    sched = LRSchedule(args, 0.008)
    t1, t2 = _Trainer(), _Trainer()
    for phase, (trainer, frac) in enumerate(((t1, 2.0 / 3.0), (t2, 1.0 / 3.0))):
        sched.bind(trainer if phase == 0 else None)
        loss = 1.0
        stopped_ref = stopped = False
        for step, prior, passes, last in epoch_schedule(nv, bs, ss, 90 * frac):
            loss = loss * (0.97 if rng.random() < 0.4 else 1.02)         # noisy plateau
            r = bool(ref.decay_learning_rate(prior, passes, loss))
            s = sched.update(prior, passes, lambda: loss)
            assert r == s
            assert abs(sched.lr - opt.param_groups[0]['lr']) <= 1e-12 * opt.param_groups[0]['lr']
            if r:
                stopped_ref = stopped = True
                break
        if phase == 0:
            assert t1.lr is None or abs(t1.lr - sched.lr) < 1e-15
        else:
            assert t2.lr is None                # phase 2: the strategy is still bound to the phase-1 optimiser
    assert sched.decays >= 2
    if smallify_decay:
        assert stopped_ref and stopped          # the plateau strategy ends the run once lr <= 1e-7
