"""The ATen-op CPU port used for baseline timing agrees with the (golden-pinned) numpy oracle."""
import numpy as np
import torch

from oracle import fvsrn_numpy as O
from oracle import torch_port as TP


def test_port_forward_and_grads_match_oracle():
    spec = O.Spec(8, 15, 32, 4, 2, 'db2', '')
    sd = TP.make_state(spec, seed=3)
    port = TP.CpuPort(spec, sd)
    rng = np.random.default_rng(0)
    coords = (rng.uniform(-1, 1, size=(64, 3))).astype(np.float32)
    w = rng.normal(size=(64, 1)).astype(np.float32)
    y = port.forward(torch.from_numpy(coords))
    (y * torch.from_numpy(w)).sum().backward()
    yo, ctx = O.model_forward(sd, spec, coords, training=True, keep=True)
    go = O.model_backward(w, ctx, spec)
    assert np.abs(y.detach().numpy() - yo).max() < 1e-5 * np.abs(yo).max()
    for k, p in port.params.items():
        assert np.abs(p.grad.numpy() - go[k]).max() <= 1e-5 * np.abs(go[k]).max(), k


def test_port_train_step_runs_and_learns():
    spec = O.Spec(4, 15, 32, 4, 2, 'db2', '')
    port = TP.CpuPort(spec, TP.make_state(spec, 1))
    xs = torch.linspace(-1, 1, 20)
    vol = (torch.sin(3 * xs)[:, None, None] * torch.cos(2 * xs)[None, :, None] * xs[None, None, :]).contiguous()
    gen = torch.Generator().manual_seed(0)
    losses = [port.train_step(vol, 512, gen) for _ in range(30)]
    assert losses[-1] < losses[0]
