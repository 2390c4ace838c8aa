import numpy as np
import torch

TOL = 1e-5      # normwise max|a-b| / max|b| for fp32 FMA paths (BASELINE.json north_star)


def nerr(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def build_pair(tabs, in_ch, chans, latent, pre_sig, is_vae, seed, device, bias_scale=0.05):
    """(oracle Net, oracle params on CPU, sdvae_b200 Model on `device` with the same weights)."""
    from oracle import sdvae_oracle as orc
    from sdvae_b200.model import Model
    sp, dn, up = tabs.spiral_tensors(), tabs.down_tensors(), tabs.up_tensors()
    net = orc.Net(in_ch, chans, latent, sp, dn, up, pre_sig, is_vae)
    params = orc.xavier_params(net.param_shapes(), seed=seed, bias_scale=bias_scale)
    model = Model(in_ch, chans, latent, [s.to(device) for s in sp], [d.to(device) for d in dn],
                  [u.to(device) for u in up], pre_sig, is_vae).to(device)
    model.load_state_dict({k: v.to(device) for k, v in params.items()}, strict=True)
    return net, params, model


def rand(shape, seed, scale=1.0):
    return torch.from_numpy((np.random.RandomState(seed).randn(*shape) * scale).astype(np.float32))
