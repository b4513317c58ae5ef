"""Print forward / gradient errors of every golden model case (and the fused step vs oracle) for the loaded library."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.util import MODEL_CASES, build_model, noise_of, relerr, replay_noise
for tag in MODEL_CASES:
    model, g, cfg = build_model(tag)
    coords = torch.from_numpy(g['coords']).cuda().requires_grad_(True)
    wout = torch.from_numpy(g['wout']).cuda()
    with replay_noise(noise_of(g)):
        y = model(coords)
    ef = relerr(y.detach().cpu().numpy(), g['y_train'])
    eg = 0.0
    if 'h64' not in tag:
        (y * wout).sum().backward()
        for name, prm in model.named_parameters():
            ref = g['grad.' + name]
            if ref.size == 0 or prm.grad is None:
                continue
            eg = max(eg, float(np.abs(prm.grad.cpu().numpy() - ref).max()) / max(float(np.abs(ref).max()), 1e-30))
    model.eval()
    tile = torch.from_numpy(g['tile']).cuda()
    with torch.no_grad(), replay_noise(noise_of(g, 'noise_eval')):
        ye = model(tile)
    ee = relerr(ye.cpu().numpy(), g['y_eval'])
    print('%-32s fwd %.2e  eval %.2e  grad %.2e' % (tag, ef, ee, eg))
