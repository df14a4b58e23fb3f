"""ORACLE (test infrastructure, not product code): PyTorch-fp32 restatement of the reference's
projection discriminator (disc.py:8-38, nets.py:26-33, torch's hook-based spectral_norm) and of
one G+D training iteration of the class-conditioned trainer in its supervised branch
(t_cls_train.py:288-312 D update, :226-286 G update, :184-185 optimisers).

The weather estimator (a torchvision ResNet-101 loaded from a private checkpoint,
t_cls_train.py:172-177) is third-party and absent, so the term g_loss_w (t_cls_train.py:256) is
left out: g_loss = g_loss_adv + loss_con.  Everything else follows the reference step by step.

Parity status: discriminator forward (incl. the in-place power iteration) PINNED bit-exact against
the live reference SNDisc (oracle/pin_train_against_reference.py); the step itself is a
restatement driven with the reference's own modules in that script (the trainer cannot be
imported: it parses argv and opens private NAS paths at import time).
"""
import torch
import torch.nn.functional as F

from . import cunet_oracle as G


# ---------------------------------------------------------------------------- discriminator
def _sn_weight(sd, name, train):
    """torch.nn.utils.spectral_norm (1 power iteration per training forward, eps 1e-12): the
    buffers `<name>_u`, `<name>_v` are updated in place, sigma = u^T W v, weight = W / sigma."""
    w = sd[f"{name}_orig"]
    u, v = sd[f"{name}_u"], sd[f"{name}_v"]
    mat = w.reshape(w.shape[0], -1)
    if train:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(mat.t(), u), dim=0, eps=1e-12))
            u.copy_(F.normalize(torch.mv(mat, v), dim=0, eps=1e-12))
        u, v = u.clone(), v.clone()
    sigma = torch.dot(u, torch.mv(mat, v))
    return w / sigma


def disc_forward(sd, x, c, train=True):
    """disc.py:27-38: 4 x (SN-conv3x3 -> SN-conv3x3 stride 2 -> LeakyReLU 0.2), global SUM pool,
    SN-linear + projection.  Returns [out, c1, c2, c3, c4]."""
    feats, h = [], x
    for i in range(1, 5):
        h = F.conv2d(h, _sn_weight(sd, f"conv{i}.0.weight", train), sd[f"conv{i}.0.bias"], padding=1)
        h = F.conv2d(h, _sn_weight(sd, f"conv{i}.1.weight", train), sd[f"conv{i}.1.bias"], padding=1,
                     stride=2)
        h = F.leaky_relu(h, 0.2)
        feats.append(h)
    pooled = feats[-1].sum(dim=(2, 3))
    out = F.linear(pooled, _sn_weight(sd, "l.weight", train), sd["l.bias"])
    proj = F.linear(c, _sn_weight(sd, "embed.weight", train), sd["embed.bias"])
    out = out + (proj * pooled).sum(dim=1, keepdim=True)
    return [out] + feats


def disc_conv_flops(H, W, B=1):
    total, h, w = 0, H, W
    for cin, cout in ((3, 64), (64, 128), (128, 256), (256, 512)):
        total += 2 * 9 * cin * cin * h * w
        h, w = h // 2, w // 2
        total += 2 * 9 * cin * cout * h * w
    return total * B


# ---------------------------------------------------------------------------- one iteration
class Trainer:
    """State of the restated trainer: leaf parameter dicts, buffers and the two Adam optimisers
    (lr, betas=(0, 0.999), weight_decay=lr/20 — t_cls_train.py:184-185)."""

    def __init__(self, g_sd, d_sd, lr=1e-4):
        self.g = {k: v.detach().clone().requires_grad_(True)
                  for k, v in g_sd.items()}
        self.d = {}
        for k, v in d_sd.items():
            t = v.detach().clone()
            is_param = k.endswith("_orig") or k.endswith("bias")
            self.d[k] = t.requires_grad_(True) if is_param else t
        self.g_opt = torch.optim.Adam(list(self.g.values()), lr=lr, betas=(0.0, 0.999),
                                      weight_decay=lr / 20)
        self.d_opt = torch.optim.Adam([v for v in self.d.values() if v.requires_grad], lr=lr,
                                      betas=(0.0, 0.999), weight_decay=lr / 20)

    def step(self, images, c_real, c_target, masks_d=None, masks_g=None, train=True,
             estimator=None, eps_con=1e-2):
        """One iteration: D update then G update.  Returns dict of the losses.
        masks_*: optional injected dropout masks for the generator forward of each update.
        estimator / eps_con: the estimator-conditioned trainer (t_est_train.py:214-283) — a frozen
        module (B,3,H,W)->(B,nc) adds g_loss_w = MSE(estimator(fake), target) (:232,:237,
        ops.py:37-39) and the reconstruction weight uses eps 1e-7 (:242); c_real is then
        estimator(images).detach() (:219,:266-267), computed by the caller."""
        # --- D update (t_cls_train.py:288-312)
        self.d_opt.zero_grad()
        real = disc_forward(self.d, images, c_real, train)[0]
        fake_img = G.forward(self.g, images, c_target, train=train, masks=masks_d)
        fake = disc_forward(self.d, fake_img.detach(), c_target, train)[0]
        d_loss = F.relu(1. - real).mean() + F.relu(1. + fake).mean()      # ops.py:42-45
        d_loss.backward()
        self.d_opt.step()
        # --- G update (t_cls_train.py:226-286, supervised branch)
        self.g_opt.zero_grad()
        fake_img = G.forward(self.g, images, c_target, train=train, masks=masks_g)
        fake = disc_forward(self.d, fake_img, c_target, train)[0]
        g_adv = (-fake).mean()                                            # ops.py:47-48
        g_l1 = F.l1_loss(fake_img, images)                                # logged only (:255)
        diff = (fake_img - images).abs().mean(dim=(1, 2, 3))              # :259-262
        lmda = (c_real - c_target).abs().mean(dim=1)
        loss_con = (diff / (lmda + eps_con)).mean()
        g_loss = g_adv + loss_con
        g_w = None
        if estimator is not None:
            g_w = F.mse_loss(estimator(fake_img), c_target)               # t_est_train.py:232,237
            g_loss = g_loss + g_w
        g_loss.backward()
        self.g_opt.step()
        out = {"d_loss": d_loss.item(), "g_loss": g_loss.item(), "g_loss_adv": g_adv.item(),
               "g_loss_l1": g_l1.item(), "loss_con": loss_con.item()}
        if g_w is not None:
            out["g_loss_w"] = g_w.item()
        return out


def step_flops(H, W, B=1):
    """Convolution FLOPs of the reference-faithful iteration: 2 G fwd + 1 G bwd + 3 D fwd + 3 D bwd
    (388.8 GF/img at 256x256)."""
    gf = G.conv_flops(H, W)
    first = 2 * 9 * 3 * 64 * H * W
    df = disc_conv_flops(H, W)
    return B * (2 * gf + (2 * gf - first) + 3 * df + 3 * (2 * df))
