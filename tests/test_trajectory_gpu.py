"""Multi-step training parity on the GPU (N = 50 fixed-seed iterations).

What these tests exist to catch: the reference's g_opt.step() (t_cls_train.py:273) changes the
weights the next inference(images, labels) (:302, :242) runs on.  The tensor-core convolutions here
read derived bf16 copies of the fp32 masters, so an optimiser that writes through raw pointers can
leave them stale (round 1 shipped exactly that).  Checked: (i) after k fused-Adam steps at a large
lr the module's output is bit-identical to a fresh module that loads the same masters; (ii) the
copies the Adam launch emits equal wu_pack_conv3x3_weights of the masters bit for bit; (iii) over
50 iterations the parameters of G and D move the way the oracle trainer's do (which is pinned
bit-for-bit to the reference modules, oracle/pin_train50_against_reference.py), and a run with
frozen operand copies is REJECTED by the same criteria (negative control)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD50 = os.path.join(os.path.dirname(__file__), "golden", "train50_b2_h32_seed0.npz")


def _mk(cuda, nc=5):
    from weather_unet_b200 import Conditional_UNet
    from weather_unet_b200.disc import SNDisc
    torch.manual_seed(0)
    G = Conditional_UNet(nc).to(cuda).train()
    torch.manual_seed(100)
    D = SNDisc(nc).to(cuda).train()
    return G, D


def test_fused_adam_emits_packed_weights(cuda):
    """wu_adam_multi with packed records: masters == torch.optim.Adam's, and w_fprop / w_dgrad ==
    wu_pack_conv3x3_weights(master) bit for bit, for every generator shape (cin 64..768)."""
    from weather_unet_b200 import Conditional_UNet, _ops as K
    from weather_unet_b200.optim import FusedAdam
    torch.manual_seed(0)
    G = Conditional_UNet(5).to(cuda)
    torch.manual_seed(0)
    R = Conditional_UNet(5).to(cuda)
    lr = 1e-2
    oa = FusedAdam(G.parameters(), lr=lr, betas=(0.0, 0.999), weight_decay=lr / 20).attach_packed(G)
    ob = torch.optim.Adam(R.parameters(), lr=lr, betas=(0.0, 0.999), weight_decay=lr / 20)
    gen = torch.Generator().manual_seed(5)
    for it in range(3):
        for (n, p), (_, q) in zip(G.named_parameters(), R.named_parameters()):
            if n.endswith("emb.weight"):
                continue
            p.grad = (torch.randn(p.shape, generator=gen) * 0.1).to(cuda)
            q.grad = p.grad.clone()
        v0 = {n: p._version for n, p in G.named_parameters()}
        oa.step()
        ob.step()
        for n, p in G.named_parameters():
            if p.grad is not None:
                assert p._version > v0[n], f"{n}: version counter not bumped by the raw-pointer update"
    for (n, p), (_, q) in zip(G.named_parameters(), R.named_parameters()):
        assert torch.allclose(p, q, rtol=2e-6, atol=2e-7), n
    names = G.packed_weight_names()
    assert len(names) == 13
    for n in names:
        w = G.get_parameter(n).detach()
        wf, wd = G._packed.buffers(n, w)
        rf, rd = K.pack_conv3x3_weights(w)
        assert torch.equal(wf, rf) and torch.equal(wd, rd), n
        # and the cache considers them current: no repack launch on the next get()
        n0 = K.launch_count()
        G._packed.get(n, G.get_parameter(n))
        assert K.launch_count() == n0, n
    # state_dict is torch.optim.Adam's: step / exp_avg / exp_avg_sq per parameter, no device tables
    sd = oa.state_dict()
    assert set(next(iter(sd["state"].values())).keys()) == {"step", "exp_avg", "exp_avg_sq"}
    assert all(set(g.keys()) == set(ob.state_dict()["param_groups"][0].keys()) for g in sd["param_groups"])
    ob2 = torch.optim.Adam(R.parameters(), lr=lr, betas=(0.0, 0.999), weight_decay=lr / 20)
    ob2.load_state_dict(sd)  # interchangeable with torch's optimiser


@pytest.mark.parametrize("fused", [True, False])
def test_generator_sees_optimizer_updates(cuda, fused):
    """After k GDTrainStep iterations at a large lr, G(x, c, seed) is bit-identical to a fresh module
    that load_state_dict()s G's masters (i.e. the operand copies are never stale), with the fused
    optimiser and with torch.optim.Adam."""
    from weather_unet_b200 import Conditional_UNet, _ops as K
    from weather_unet_b200.train_step import GDTrainStep
    G, D = _mk(cuda)
    step = GDTrainStep(G, D, lr=5e-3, fused_adam=fused)
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(2, 3, 32, 32, generator=g) * 2 - 1).to(cuda)
    cr = torch.eye(5)[torch.randint(0, 5, (2,), generator=g)].to(cuda)
    ct = torch.eye(5)[torch.randint(0, 5, (2,), generator=g)].to(cuda)
    with torch.no_grad():
        y0 = G(x, ct, seed=7).clone()
    for it in range(3):
        n0 = K.launch_count()
        step.step(x, cr, ct)
        if it > 0:
            launches = K.launch_count() - n0
    with torch.no_grad():
        y1 = G(x, ct, seed=7)
    fresh = Conditional_UNet(5).to(cuda).train()
    fresh.load_state_dict(G.state_dict())
    with torch.no_grad():
        y2 = fresh(x, ct, seed=7)
    assert torch.equal(y1, y2), (y1 - y2).abs().max().item()
    assert (y1 - y0).abs().max().item() > 1e-2, "three Adam steps at lr 5e-3 must change the output"
    print(f"fused={fused}: {launches} library launches per iteration")
    # a second optimiser step with the saved graph of an earlier forward must be refused by autograd
    y = G(x, ct, seed=7)
    step.g_opt.zero_grad(set_to_none=True)
    y.mean().backward(retain_graph=True)
    step.g_opt.step()
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        y.mean().backward()


def _run50(cuda, z, runner):
    """Drive `runner(images, c_real, c_target, masks_d, masks_g) -> dict of losses` for the 50
    golden iterations with the golden mask stream."""
    from oracle.pin_train50_against_reference import draw_masks
    x, cr, ct = (torch.from_numpy(z[k]).to(cuda) for k in ("images", "c_real", "c_target"))
    keys = [str(k) for k in z["keys"]]
    B, H = x.shape[0], x.shape[2]
    torch.manual_seed(int(z["seed_masks"][0]))
    curve = []
    for it in range(z["curve"].shape[0]):
        md = tuple(m.to(cuda) for m in draw_masks(B, H, H))
        mg = tuple(m.to(cuda) for m in draw_masks(B, H, H))
        out = runner(x, cr, ct, md, mg)
        curve.append([float(out[k]) for k in keys])
    return np.array(curve), keys


def _movement(final, init):
    """Per tensor: (cosine between the two displacement vectors, |d_a - d_b| / |d_b|) helpers."""
    return {k: (final[k].detach().double() - init[k].double()) for k in init}


def test_train50_parameter_trajectory(cuda):
    from oracle import train_oracle as T
    from weather_unet_b200.train_step import GDTrainStep
    z = np.load(GOLD50)
    lr = float(z["lr"][0])
    gold = z["curve"]

    # --- the oracle trainer on the GPU in fp32 (TF32 off): ties the GPU run to the pinned CPU curve
    G, D = _mk(cuda)
    g0 = {k: v.detach().clone() for k, v in G.state_dict().items()}
    d0 = {k: v.detach().clone() for k, v in D.state_dict().items()}
    orc = T.Trainer(g0, d0, lr=lr)
    c_orc, keys = _run50(cuda, z, lambda x, cr, ct, md, mg: orc.step(x, cr, ct, masks_d=md, masks_g=mg))
    ki = {k: i for i, k in enumerate(keys)}
    e_first = np.abs(c_orc[:5] - gold[:5]).max()
    print("oracle on GPU vs golden (CPU reference modules), first 5 iterations: max |diff| =", e_first)
    assert e_first < 5e-3
    mv_orc_g = _movement(orc.g, g0)
    mv_orc_d = _movement({k: v for k, v in orc.d.items()}, d0)
    # the pinned CPU run and the GPU oracle run move every tensor by the same amount (the dynamics
    # are chaotic, so direction is only compared between runs on the same machine below)
    for name, st in zip([str(n) for n in z["g_names"]], z["g_stats"]):
        if name.endswith("emb.weight"):
            continue
        m = mv_orc_g[name].norm().item()
        assert abs(m - st[1]) <= 0.25 * st[1] + 1e-6, (name, m, st[1])

    def run_ours(stale):
        G, D = _mk(cuda)
        step = GDTrainStep(G, D, lr=lr)
        if stale:  # negative control: what round 1 did — operand copies frozen at their first value
            step.g_opt._packed_of = {}
            G._packed.get = (lambda orig: (lambda name, w: (orig(name, w) if name not in G._packed._key
                                                            else G._packed._buf[name])))(G._packed.get)
        curve, _ = _run50(cuda, z, lambda x, cr, ct, md, mg: step.step(x, cr, ct, masks_d=md, masks_g=mg))
        return curve, _movement(dict(G.state_dict()), g0), _movement(dict(D.state_dict()), d0)

    def criteria(curve, mv_g, mv_d):
        """-> (worst cosine over G's 3x3 weights, median cosine over all G tensors, relative error of
        the last-10-iteration mean of loss_con, median cosine over D tensors)."""
        cos_g = {}
        for k, a in mv_g.items():
            b = mv_orc_g[k]
            if b.norm().item() == 0:
                continue
            cos_g[k] = (a.flatten() @ b.flatten() / (a.norm() * b.norm() + 1e-30)).item()
        cos_d = []
        for k, a in mv_d.items():
            b = mv_orc_d[k]
            if k.endswith("_u") or k.endswith("_v") or b.norm().item() == 0:
                continue
            cos_d.append((a.flatten() @ b.flatten() / (a.norm() * b.norm() + 1e-30)).item())
        conv = [v for k, v in cos_g.items() if k.endswith(".weight") and k.startswith("dconv")]
        tail = lambda c: c[-10:, ki["loss_con"]].mean()
        return (min(conv), float(np.median(list(cos_g.values()))),
                abs(tail(curve) - tail(c_orc)) / tail(c_orc), float(np.median(cos_d)), cos_g)

    curve, mv_g, mv_d = run_ours(stale=False)
    worst_conv, med_g, tail_err, med_d, cos_g = criteria(curve, mv_g, mv_d)
    print("ours vs oracle (same GPU): worst cos over G conv weights %.4f, median cos G %.4f, "
          "loss_con tail error %.3f, median cos D %.4f" % (worst_conv, med_g, tail_err, med_d))
    print("loss_con every 7 iterations: ours", np.round(curve[::7, ki["loss_con"]], 3),
          "oracle", np.round(c_orc[::7, ki["loss_con"]], 3), "golden", np.round(gold[::7, ki["loss_con"]], 3))
    for k, v in sorted(cos_g.items(), key=lambda kv: kv[1])[:6]:
        print("   lowest cosines:", k, round(v, 4))
    # the curve really moves (a frozen generator cannot follow it) ...
    assert gold[-10:, ki["loss_con"]].mean() < 0.7 * gold[:3, ki["loss_con"]].mean()
    assert curve[-10:, ki["loss_con"]].mean() < 0.7 * curve[:3, ki["loss_con"]].mean()
    # ... per-iteration losses follow the oracle while the two runs are still on one trajectory
    err = np.abs(curve - c_orc) / np.maximum(np.abs(c_orc), 1.0)
    big = [ki[k] for k in ("g_loss", "loss_con", "g_loss_l1")]
    print("per-iteration relative error (g_loss, loss_con, g_loss_l1), every 5:\n",
          np.array2string(err[::5][:, big], precision=3))
    assert err[:10][:, big].max() < 5e-2, err[:10]
    assert tail_err < 0.15
    # ... and 50 updates displaced every tensor in the oracle's direction
    assert worst_conv > 0.8 and med_g > 0.9 and med_d > 0.8, (worst_conv, med_g, med_d)

    # negative control: with stale operand copies the same criteria must FAIL
    curve_s, mv_gs, mv_ds = run_ours(stale=True)
    worst_s, med_s, tail_s, med_ds, _ = criteria(curve_s, mv_gs, mv_ds)
    print("stale copies: worst cos %.4f, median cos G %.4f, tail error %.3f" % (worst_s, med_s, tail_s))
    assert not (worst_s > 0.8 and med_s > 0.9 and tail_s < 0.15), "the criteria cannot see frozen weights"
