"""Final-PSNR parity with the reference after the same two-phase training run (BASELINE.json north_star).
The golden values come from running the reference's own training() in the build container
(tests/golden/make_psnr_golden.py): same torch seed, same DataLoader sample stream, same schedule.  Gate: 0.05 dB
is the north-star figure; fp32 re-association makes the runs drift chaotically, yet the measured delta on B200 is 0.0000 dB for both runs; the asserted gate is the north-star 0.05 dB
(the reference itself spreads 0.8-1.3 dB between seeds) and the measured delta is printed."""
import ast
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _args(g):
    return {k: ast.literal_eval(v) for k, v in zip(g['args_keys'].tolist(), g['args_vals'].tolist())}


@pytest.mark.parametrize('tag', ['basic', 'smallify'])
def test_final_psnr_matches_reference_run(tag):
    from tests import ref_loop
    g = np.load(os.path.join(GOLD, 'psnr_run_%s.npz' % tag))
    args = _args(g)
    torch.manual_seed(0)
    info = ref_loop.training(args, g['volume_raw'])
    delta = info['psnr'] - float(g['psnr'])
    print('\n[psnr parity] %s: reference %.4f dB, this repo %.4f dB, delta %+.4f dB, zeros %s vs %s, steps %d' % (
        tag, float(g['psnr']), info['psnr'], delta, info['num_zeros'], float(g['num_zeros']), info['steps']))
    assert info['num_parameters'] == int(g['num_parameters'])
    assert abs(delta) < 0.05


def test_fast_loop_reaches_the_same_quality():
    """The graph-captured loop uses its own (Philox) sample stream, so only the quality level is comparable."""
    from latent_feature_grid_compression_b200.data.IndexDataset import normalize_volume
    from latent_feature_grid_compression_b200.training.fast_loop import train_volume
    g = np.load(os.path.join(GOLD, 'psnr_run_basic.npz'))
    args = _args(g)
    vol = torch.from_numpy(g['volume_raw'])
    vol = normalize_volume(vol, torch.min(vol), torch.max(vol), -1.0, 1.0)
    torch.manual_seed(0)
    info = train_volume(args, volume=vol, seed=3)
    print('\n[psnr fast loop] reference %.3f dB, fast loop %.3f dB' % (float(g['psnr']), info['psnr']))
    assert info['psnr'] > float(g['psnr']) - 1.5


def test_fast_loop_psnr_distribution_matches_the_reference_at_a_baseline_config():
    """BASELINE config turbulence_basic (150^3, no pruning, the full 50-pass two-phase schedule): the fast loop's final
    PSNR over three Philox seeds against the reference's own three runs on the same synthetic volume
    (tests/golden/psnr_configs.json, unmodified training/training.py:184 on CPU).  The sample streams differ by
    construction, so the gate is distributional: mean within 0.3 dB (the reference's own seed spread is 0.23 dB)."""
    import json
    import bench
    from latent_feature_grid_compression_b200.training.fast_loop import train_volume
    recs = [r for r in json.load(open(os.path.join(GOLD, 'psnr_configs.json')))
            if r['config'] == 'turbulence_basic' and r['max_pass'] == r['config_max_pass']]
    assert len(recs) >= 3
    vol = bench.synthetic_volume(150, 'cuda').cpu()
    mine = []
    for s in range(3):
        torch.manual_seed(s)
        info = train_volume(dict(recs[0]['args']), volume=vol, seed=1000 + s)
        assert info['steps'] == recs[-1]['optimiser_steps']          # same schedule: same number of optimiser steps
        mine.append(info['psnr'])
    ref = [r['psnr'] for r in recs]
    print('\n[psnr @ turbulence_basic] reference %s, fast loop %s' % (['%.2f' % v for v in ref], ['%.2f' % v for v in mine]))
    assert abs(np.mean(mine) - np.mean(ref)) < 0.3


def test_fast_loop_psnr_distribution_on_the_variational_config():
    """BASELINE config mhd_p_dynamic_variational (255^3, variational dropout with the learned variance model) at 12 of
    its 60 passes: six fast-loop seeds against the reference's three runs.  One and the same seed spreads ~0.2 dB between
    runs here (the scatter atomics reorder the sums and the noisy objective amplifies that), so single runs say little:
    24 runs measured on B200 gave 47.57 +- 0.19 dB against the reference's 47.63 +- 0.11
    (profiles/r2_variational_psnr_distribution.txt).  Gate: means within 0.3 dB (three standard errors)."""
    import json
    import bench
    from latent_feature_grid_compression_b200.training.fast_loop import train_volume
    recs = [r for r in json.load(open(os.path.join(GOLD, 'psnr_configs.json')))
            if r['config'] == 'mhd_p_dynamic_variational' and r['max_pass'] == 12]
    assert len(recs) >= 3
    vol = bench.synthetic_volume(255, 'cuda').cpu()
    mine = []
    for s in range(6):
        torch.manual_seed(s)
        info = train_volume(dict(recs[0]['args']), volume=vol, seed=1000 + s)
        assert info['steps'] == recs[0]['optimiser_steps']
        mine.append(info['psnr'])
    ref = [r['psnr'] for r in recs]
    print('\n[psnr @ mhd_p_dynamic_variational, 12 passes] reference %s, fast loop %s' % (
        ['%.2f' % v for v in ref], ['%.2f' % v for v in mine]))
    assert abs(np.mean(mine) - np.mean(ref)) < 0.3
