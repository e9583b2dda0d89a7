"""Reconstruction-quality report on the device: the step *after* the hot path.

``report`` mirrors the vision branch of ``breaching.analysis.report`` (``analysis/analysis.py:14-115, 204-283``) for the
metrics that need no third-party network or dataset: per-example MSE / PSNR of the de-normalised, clamped batches
(``bre_image_mse`` -- ``analysis.py:228-242``, ``metrics.py:108-130``), feature MSE on the attacked model (engine forward,
``analysis.py:56-69``), label accuracy (``count_integer_overlap`` ``:286-316``).  LPIPS, CW-SSIM, registered PSNR and IIP
(lpips / kornia / dataset dependent) are reported as NaN, never silently approximated.
"""
import math

import torch

from . import engine as E


def psnr_compute(img_batch, ref_batch, batched=False, factor=1.0, clip=False):
    """analysis/metrics.py:108-130 on the device: (mean PSNR, max PSNR) over the examples, or the batch PSNR."""
    mse = E.image_mse(img_batch, ref_batch, clamp=clip)
    if batched:
        m = sum(mse) / len(mse)
        if m > 0 and math.isfinite(m):
            return 10 * math.log10(factor ** 2 / m)
        return [float("nan")] * 2 if not math.isfinite(m) else [float("inf")] * 2
    if any(m == 0 for m in mse):
        return [float("inf")] * 2
    if not all(math.isfinite(m) for m in mse):
        return [float("nan")] * 2
    per_example = [10 * math.log10(factor ** 2 / m) for m in mse]
    return sum(per_example) / len(per_example), max(per_example)


def count_integer_overlap(rec_labels, true_labels, maxlength):
    """analysis/analysis.py:286-316."""
    if rec_labels is None:
        return 0
    a = torch.bincount(rec_labels.view(-1), minlength=maxlength)
    b = torch.bincount(true_labels[true_labels != -100].view(-1), minlength=maxlength)
    return float(1 - (a - b).abs().sum() / 2 / rec_labels.numel())


def report(reconstructed_user_data, true_user_data, server_payload, model_template=None, setup=None, **unused):
    """Vision metrics of one attack: ``dict(mse, psnr, max_psnr, feat_mse, label_acc, parameters, lpips, rpsnr, ssim, ...)``."""
    metadata = server_payload[0]["metadata"]
    if getattr(metadata, "modality", "vision") != "vision":
        raise NotImplementedError("text metrics need tokenizers / datasets: not part of the engine")
    dev = torch.device(setup["device"]) if setup is not None else reconstructed_user_data["data"].device
    rec = reconstructed_user_data["data"].to(dev, torch.float32)
    ref = true_user_data["data"].to(dev, torch.float32)
    mean = getattr(metadata, "mean", None)
    std = getattr(metadata, "std", None)
    mse = E.image_mse(rec, ref, mean, std, clamp=True)                    # analysis.py:228-236
    psnr = [10 * math.log10(1.0 / m) if m > 0 else float("inf") for m in mse]
    out = dict(mse=sum(mse) / len(mse), max_mse=max(mse), psnr=sum(psnr) / len(psnr), max_psnr=max(psnr),
               lpips=float("nan"), rpsnr=float("nan"), ssim=float("nan"), max_ssim=float("nan"), max_rpsnr=float("nan"), order=None)
    labels = reconstructed_user_data.get("labels")
    classes = getattr(metadata, "classes", None) or (int(true_user_data["labels"].max()) + 1)
    out["label_acc"] = count_integer_overlap(labels.to(dev) if labels is not None else None, true_user_data["labels"].to(dev), classes)
    out["feat_mse"] = float("nan")
    if model_template is not None:
        import copy

        from .config import get_attack_config

        feat_mse = 0.0
        for payload in server_payload:                                       # analysis.py:56-69
            m = copy.deepcopy(model_template).to(dev)
            buffers = payload["buffers"] if payload["buffers"] is not None else true_user_data.get("buffers")
            with torch.no_grad():
                for p, src in zip(m.parameters(), payload["parameters"]):
                    p.copy_(src.to(dev))
                if buffers:
                    for b, src in zip(m.buffers(), buffers):
                        b.copy_(src.to(dev))
            m.eval()
            eng = E.Engine(m, tuple(rec.shape), get_attack_config("invertinggradients"), dev, backend="simt")   # metric: fp32 forward
            eng.load_model()
            fr, ft = eng.forward(rec), eng.forward(ref)
            eng.close()
            rel = true_user_data["labels"].view(-1).to(dev)
            feat_mse += float((fr - ft)[..., rel].pow(2).mean())
        out["feat_mse"] = feat_mse
        out["parameters"] = sum(p.numel() for p in model_template.parameters())
    return out
