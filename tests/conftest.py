import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 (B200) device; run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_files():
    """Module-level fixtures (reference module run + captured SDPA calls); the prepare_* fixtures are listed apart."""
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".pt") and not f.startswith(("prepare", "cross_", "site_")))


def cross_golden_files():
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".pt") and f.startswith("cross_"))


def prepare_golden_files():
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".pt") and f.startswith("prepare_"))


def prepare_seq_golden_files():
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".pt") and f.startswith("prepareseq_"))


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def make_qkv(N, Tq, Tk, H, G, hd, seed, unit_norm=True, device="cpu", dtype=torch.bfloat16):
    """Synthetic inputs of SURVEY.md §8(d): q,k ~ N(0,1) L2-normalised over hd (mirrors apply_qk_norm), v ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(N, Tq, H, hd, generator=g)
    k = torch.randn(N, Tk, G, hd, generator=g)
    v = torch.randn(N, Tk, G, hd, generator=g)
    if unit_norm:
        q = torch.nn.functional.normalize(q, dim=-1)
        k = torch.nn.functional.normalize(k, dim=-1)
    return q.to(dtype).to(device), k.to(dtype).to(device), v.to(dtype).to(device)
