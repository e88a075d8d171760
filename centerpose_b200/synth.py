"""Seeded synthetic inputs for the CenterPose hot path.

There are no checkpoints and no dataset on the build or GPU boxes, so every
test / benchmark input is generated here, bit-identically on any machine:

* `seeded_state_dict`  -- random weights with the reference's parameter names
  and shapes (BN statistics perturbed so folding is observable, DCN offset
  convs non-zero so the deformable sampling actually deforms).
* `planted_heads`      -- head tensors (logits) that contain a known set of
  projected cuboids: Gaussian peaks in `hm` / `hm_hp`, exact `reg`, `wh`,
  `hps`, `hp_offset`, `scale` at the peak cells, plus sub-threshold jitter
  that breaks top-K ties (SURVEY.md 8d "tie hazard").
* `synthetic_frames`   -- uint8 Objectron-shaped RGB frames.
"""
import math

import numpy as np
import torch

DEFAULT_HEADS = {"hm": 1, "wh": 2, "hps": 16, "reg": 2, "hm_hp": 8, "hp_offset": 2, "scale": 3}
TRACKING_HEADS = {"hm": 1, "wh": 2, "hps": 16, "hps_uncertainty": 16, "reg": 2, "hm_hp": 8,
                  "hp_offset": 2, "scale": 3, "scale_uncertainty": 3, "tracking": 2, "tracking_hp": 16}


def seeded_state_dict(model, seed=0, offset_std=1.0, head_gain=1.0):
    """Deterministic fp32 values for every entry of `model.state_dict()`.

    Each tensor has its own generator seeded from (seed, key) so the result
    does not depend on parameter iteration order."""
    out = {}
    sd = model.state_dict()
    for k in sd:
        v = sd[k]
        h = 1469598103934665603
        for ch in (k + "#%d" % seed).encode():
            h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        g = torch.Generator().manual_seed(h & 0x7FFFFFFFFFFFFFFF)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros_like(v)
        elif k.endswith("running_mean"):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            out[k] = torch.rand(v.shape, generator=g) * 0.5 + 0.75
        elif v.dim() == 1 and k.endswith(".weight"):          # BN / GN gamma
            out[k] = torch.rand(v.shape, generator=g) * 0.5 + 0.75
        elif v.dim() == 1:                                    # biases, BN beta
            b = torch.randn(v.shape, generator=g) * 0.05
            if k.split(".")[0] in ("hm", "hm_hp") and not k.endswith(".0.bias"):
                b = b - 2.19
            out[k] = b
        elif v.dim() == 4 and ".up_" in k:                    # depthwise ConvTranspose, bilinear-ish
            f = v.shape[2] // 2
            c = (2 * f - 1 - f % 2) / (2.0 * f)
            w = torch.zeros(v.shape[2], v.shape[3])
            for i in range(v.shape[2]):
                for j in range(v.shape[3]):
                    w[i, j] = (1 - abs(i / f - c)) * (1 - abs(j / f - c))
            out[k] = w.view(1, 1, *w.shape) * (1 + 0.1 * torch.randn(v.shape, generator=g))
        elif v.dim() == 4:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            std = math.sqrt(2.0 / fan_in)
            if "conv_offset_mask" in k:
                std = offset_std / math.sqrt(fan_in)
            elif k.split(".")[0] not in ("base", "dla_up", "ida_up", "convGRU") and v.shape[2] == 1:
                std = head_gain * math.sqrt(1.0 / fan_in)
            elif k.startswith("convGRU"):
                std = math.sqrt(1.0 / fan_in)
            out[k] = torch.randn(v.shape, generator=g) * std
        else:
            out[k] = torch.randn(v.shape, generator=g) * 0.01
        out[k] = out[k].to(v.dtype)
    return out


def synthetic_frames(batch, height=512, width=512, seed=317):
    """uint8 [B,H,W,3] i.i.d. U{0..255}; seed 317 is the reference's (opts.py:56)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(batch, height, width, 3), dtype=np.uint8)


def normalize_frames(frames_u8, mean=(0.408, 0.447, 0.470), std=(0.289, 0.274, 0.278)):
    """(x/255 - mean)/std, HWC->NCHW, fp32 -- base_detector.py:132-134 for an
    input that is already at network resolution (identity affine)."""
    x = frames_u8.astype(np.float32) / np.float32(255.0)
    x = (x - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2)).astype(np.float32)


def default_camera(width=512, height=512):
    """demo.py:143-144 intrinsics (600x800 portrait Objectron frame) rescaled to width x height."""
    fx = 663.0287679036459 * width / 600.0
    fy = 663.0287679036459 * height / 800.0
    return np.array([[fx, 0, 300.2775065104167 * width / 600.0],
                     [0, fy, 395.00066121419275 * height / 800.0],
                     [0, 0, 1.0]])


def _rand_rot(rng):
    # moderate rotations about a random axis
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = rng.uniform(0.2, 2.6)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def _cuboid(scale):
    sx, sy, sz = scale[0] / scale[1], 1.0, scale[2] / scale[1]
    v = []
    for x in (-sx / 2, sx / 2):
        for y in (-sy / 2, sy / 2):
            for z in (-sz / 2, sz / 2):
                v.append([x, y, z])
    return np.array(v)


def _logit(p):
    return np.log(p / (1.0 - p))


def planted_heads(n_obj=3, seed=0, heads=None, out_h=128, out_w=128, down=4, cam=None,
                  disagree_px=0.0, peak=0.95, sigma=1.5, drop_joints=()):
    """One image worth of head logits with `n_obj` planted cuboids.

    Returns (heads_dict {name: fp32 [C,out_h,out_w]}, truth dict).  All values
    are generated in float64 and rounded to fp32 once.  `drop_joints`: keypoint indices whose heat-map peak is NOT
    planted (the decode then reports the -10000 sentinel for them in rep_mode 4: the 4 - 5 point EPnP path)."""
    heads = dict(heads or DEFAULT_HEADS)
    rng = np.random.default_rng(seed)
    if cam is None:
        cam = default_camera(out_w * down, out_h * down)
    H, W = out_h, out_w
    out = {}
    for name, c in heads.items():
        if name in ("hm", "hm_hp"):
            out[name] = rng.uniform(1e-4, 2e-4, size=(c, H, W))        # probabilities for now
        elif name in ("hps_uncertainty", "scale_uncertainty"):
            out[name] = rng.normal(-1.0, 0.2, size=(c, H, W))
        else:
            out[name] = rng.normal(0, 0.01, size=(c, H, W))
    yy, xx = np.mgrid[0:H, 0:W]
    truth = {"R": [], "t": [], "scale": [], "kps_map": [], "ct_int": []}
    taken = []
    tries = 0
    while len(truth["R"]) < n_obj and tries < 2000:
        tries += 1
        scale = np.array([rng.uniform(0.3, 2.0), 1.0, rng.uniform(0.3, 2.0)]) * rng.uniform(0.1, 0.6)
        R = _rand_rot(rng)
        tz = rng.uniform(2.5, 7.0)
        t = np.array([rng.uniform(-0.35, 0.35) * tz, rng.uniform(-0.35, 0.35) * tz, tz])
        P = _cuboid(scale) @ R.T + t
        if np.any(P[:, 2] < 0.5):
            continue
        u = cam[0, 0] * P[:, 0] / P[:, 2] + cam[0, 2]
        v = cam[1, 1] * P[:, 1] / P[:, 2] + cam[1, 2]
        kx, ky = u / down, v / down
        if kx.min() < 6 or ky.min() < 6 or kx.max() > W - 7 or ky.max() > H - 7:
            continue
        x0, x1, y0, y1 = kx.min(), kx.max(), ky.min(), ky.max()
        cx, cy = (x0 + x1) / 2, (y0 + y1) / 2
        if any(abs(cx - a) < 10 and abs(cy - b) < 10 for a, b in taken):
            continue
        ix, iy = int(math.floor(cx)), int(math.floor(cy))
        # keypoint peak cells must be unique in hp_offset (2 shared channels)
        cells = [(int(math.floor(a)), int(math.floor(b))) for a, b in zip(kx, ky)]
        if len(set(cells)) < 8 or any(c in truth.get("_cells", set()) for c in cells):
            continue
        truth.setdefault("_cells", set()).update(cells)
        taken.append((cx, cy))
        pk = peak - 0.03 * len(truth["R"])
        g = pk * np.exp(-((xx - ix) ** 2 + (yy - iy) ** 2) / (2 * sigma * sigma))
        cls = len(truth["R"]) % out["hm"].shape[0]       # multi-class hm (opt.num_classes > 1): objects take classes in turn
        out["hm"][cls] = np.maximum(out["hm"][cls], g)
        if "reg" in out:
            out["reg"][:, iy, ix] = [cx - ix, cy - iy]
        out["wh"][:, iy, ix] = [(x1 - x0) * 1.1 + 2, (y1 - y0) * 1.1 + 2]
        if "scale" in out:
            out["scale"][:, iy, ix] = scale
        if "tracking" in out:
            out["tracking"][:, iy, ix] = rng.normal(0, 1.0, size=2)
        for j in range(8):
            dxy = rng.normal(0, disagree_px, size=2) if disagree_px > 0 else np.zeros(2)
            out["hps"][2 * j, iy, ix] = kx[j] - ix + dxy[0]
            out["hps"][2 * j + 1, iy, ix] = ky[j] - iy + dxy[1]
            if "tracking_hp" in out:
                out["tracking_hp"][2 * j:2 * j + 2, iy, ix] = rng.normal(0, 1.0, size=2)
            if "hm_hp" in out and j not in drop_joints:
                jx, jy = cells[j]
                pj = (peak - 0.02 * j) * np.exp(-((xx - jx) ** 2 + (yy - jy) ** 2) / (2 * sigma * sigma))
                out["hm_hp"][j] = np.maximum(out["hm_hp"][j], pj)
                if "hp_offset" in out:
                    out["hp_offset"][:, jy, jx] = [kx[j] - jx, ky[j] - jy]
        truth["R"].append(R)
        truth["t"].append(t)
        truth["scale"].append(scale)
        truth["kps_map"].append(np.stack([kx, ky], 1))
        truth["ct_int"].append((ix, iy))
    truth.pop("_cells", None)
    for name in ("hm", "hm_hp"):
        if name in out:
            out[name] = _logit(np.clip(out[name], 1e-6, 1 - 1e-6))
    out = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}
    truth["cam"] = cam
    return out, truth


def planted_batch(batch, n_obj=3, seed=0, heads=None, **kw):
    """Stack `planted_heads` for seeds seed..seed+batch-1: {name: [B,C,H,W]}."""
    hs, truths = [], []
    for b in range(batch):
        h, t = planted_heads(n_obj=n_obj, seed=seed + b, heads=heads, **kw)
        hs.append(h)
        truths.append(t)
    return {k: np.stack([h[k] for h in hs]) for k in hs[0]}, truths


def calibrate_head_bias(model, heads_out, target=4):
    """Random-init weights put every (or no) heat-map peak above the decode thresholds, which would make the
    decode / PnP stage do 100 (or 0) solves per frame.  Given the head logits `heads_out` of a calibration batch
    ({'hm': [B,1,h,w], 'hm_hp': [B,8,h,w]} torch tensors, any device), shift the two heat-map biases of `model` so
    that about `target` centre peaks per frame pass vis_thresh = 0.3 and about `target` peaks per keypoint channel
    pass the 0.1 gate -- a realistic scene density for the post-network stage.  Setup only (bench / tests), never
    inside a timed region.  Returns the two shifts."""
    import torch.nn.functional as F
    shifts = {}
    with torch.no_grad():
        for head, thr in (("hm", math.log(0.3 / 0.7)), ("hm_hp", math.log(0.1 / 0.9))):
            hm = heads_out[head].float()
            pk = F.max_pool2d(hm, 3, 1, 1)
            peaks = torch.where(pk == hm, hm, torch.full_like(hm, -1e9)).flatten(2)      # [B, C, HW]
            kth = peaks.topk(target + 1, dim=2).values
            mid = (0.5 * (kth[..., target - 1] + kth[..., target])).median()
            delta = float(thr - mid)
            getattr(model, head)[-1].bias += delta
            shifts[head] = delta
    return shifts
