"""CPU probe: how much Q / TD error comes from rounding which activation tensor to bf16 (i.e. dropping its `lo` plane)?"""
import sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from oracle import synth
from oracle import learner as ol
from oracle.learner import init_params, LearnerState, learner_update

torch.set_num_threads(8)
A = 9
d = synth.synthetic_batch(8, A, seed=17, ragged=True)
params = init_params(A, seed=3)
st = LearnerState(online={k: v.clone() for k, v in params.items()}, target=init_params(A, seed=4))
base = learner_update(st, synth.to_torch_batch(d), apply=False)

bf = lambda x: x.bfloat16().float()
orig_encode = ol.encode


def make_encode(drop):
    def enc(p, x):
        x = F.relu(F.conv2d(x, p["feature.0.weight"], p["feature.0.bias"], stride=4))
        if "act1" in drop: x = bf(x)
        x = F.relu(F.conv2d(x, p["feature.2.weight"], p["feature.2.bias"], stride=2))
        if "act2" in drop: x = bf(x)
        x = F.relu(F.conv2d(x, p["feature.4.weight"], p["feature.4.bias"], stride=1))
        x = x.flatten(1)
        if "act3" in drop: x = bf(x)
        x = F.relu(F.linear(x, p["feature.7.weight"], p["feature.7.bias"]))
        if "latent" in drop: x = bf(x)
        return x
    return enc


for drop in (["act1"], ["act2"], ["act3"], ["latent"], ["act1", "act2"], ["act1", "act2", "act3"], ["act1", "act2", "act3", "latent"]):
    ol.encode = make_encode(drop)
    st2 = LearnerState(online={k: v.clone() for k, v in params.items()}, target=init_params(A, seed=4))
    out = learner_update(st2, synth.to_torch_batch(d), apply=False)
    eq = (out["q"] - base["q"]).abs().max().item()
    etd = np.abs(out["td"] - base["td"]).max()
    print(f"drop lo of {'+'.join(drop):28s}: max|dq| {eq:.2e}  max|dTD| {etd:.2e}")
ol.encode = orig_encode

# ---- backward: also round the gradients flowing into the encoder GEMMs (dY operands) and the saved activations (wgrad B operands)
class RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return bf(x)

    @staticmethod
    def backward(ctx, g):
        return bf(g)


def enc_rounded(p, x):
    r = RoundGrad.apply
    x = r(F.relu(F.conv2d(x, p["feature.0.weight"], p["feature.0.bias"], stride=4)))
    x = r(F.relu(F.conv2d(x, p["feature.2.weight"], p["feature.2.bias"], stride=2)))
    x = r(F.relu(F.conv2d(x, p["feature.4.weight"], p["feature.4.bias"], stride=1))).flatten(1)
    return r(F.relu(F.linear(x, p["feature.7.weight"], p["feature.7.bias"])))


ol.encode = enc_rounded
st3 = LearnerState(online={k: v.clone() for k, v in params.items()}, target=init_params(A, seed=4))
out = learner_update(st3, synth.to_torch_batch(d), apply=True)
st0 = LearnerState(online={k: v.clone() for k, v in params.items()}, target=init_params(A, seed=4))
ol.encode = orig_encode
base2 = learner_update(st0, synth.to_torch_batch(d), apply=True)
print("balanced encoder (fwd+bwd activations/gradients bf16, weights exact):")
print(f"  max|dq| {(out['q'] - base2['q']).abs().max().item():.2e}  max|dTD| {np.abs(out['td'] - base2['td']).max():.2e}")
for k in base2["grads"]:
    g0, g1 = base2["grads"][k], out["grads"][k]
    print(f"  grad {k:28s} rel err {((g1 - g0).abs().max() / (g0.abs().max() + 1e-12)).item():.2e}   param err after Adam "
          f"{(st3.online[k] - st0.online[k]).abs().max().item():.2e}")
