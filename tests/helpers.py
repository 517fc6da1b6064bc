"""Shared helpers for the parity tests (fixture -> oracle/product inputs)."""
import numpy as np
import torch


def cell_params(arr, prefix, dtype=torch.float32):
    """prefix e.g. 'encoder.encoding_cells.0' -> oracle param dict (leaf tensors w/ grad)."""
    def t(k):
        return torch.tensor(arr[f"param:{prefix}.{k}"], dtype=dtype, requires_grad=True)
    return {"Wg": t("dconv_gate.weight"), "bg": t("dconv_gate.biases"),
            "Wc": t("dconv_candidate.weight"), "bc": t("dconv_candidate.biases")}


def cell_grads(arr, prefix):
    return {"Wg": arr[f"grad:{prefix}.dconv_gate.weight"], "bg": arr[f"grad:{prefix}.dconv_gate.biases"],
            "Wc": arr[f"grad:{prefix}.dconv_candidate.weight"],
            "bc": arr[f"grad:{prefix}.dconv_candidate.biases"]}


def supports_of(arr, dtype=torch.float32):
    out = []
    i = 0
    while f"support{i}" in arr:
        out.append(torch.tensor(arr[f"support{i}"], dtype=dtype))
        i += 1
    return out


def rel_err(a, b):
    """max|a-b| / max|b| -- the metric SURVEY 7.2 / BASELINE.json's 1e-4 refers to."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))
