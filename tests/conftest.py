import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
_TORCH_DTYPES = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _from_np(arr: np.ndarray, dtype: torch.dtype) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if dtype in (torch.float16, torch.bfloat16):
        return t.view(dtype)
    return t


class GoldenAttention:
    """One tests/golden/attn_*.npz fixture: inputs + the reference's outputs (see oracle/make_golden.py)."""

    def __init__(self, path: Path):
        z = np.load(path)
        self.name = path.stem[len("attn_"):]
        self.dtype = _TORCH_DTYPES[str(z["dtype"])]
        self.seq_lens = [tuple(int(v) for v in row) for row in z["seq_lens"]]
        self.n_qo_heads, self.n_kv_heads, self.head_dim, self.block_size, self.n_blocks = (int(v) for v in z["geometry"])
        self.fused_qkv = bool(int(z["fused_qkv"]))
        for name in ("query", "key", "value", "key_cache", "value_cache", "ref_key_cache_owned", "ref_value_cache_owned", "ref_out"):
            setattr(self, name, _from_np(z[name], self.dtype))
        self.ref_fp32 = torch.from_numpy(z["ref_fp32"])
        for name in ("q_cu_seq_lens", "kv_cu_seq_lens", "paged_kv_last_page_len", "new_cache_slots", "block_tables", "cu_blocks_lens", "owned_blocks"):
            setattr(self, name, [int(v) for v in z[name]])

    def requests(self):
        out = []
        for i, (q, kv) in enumerate(self.seq_lens):
            slots = self.new_cache_slots[self.q_cu_seq_lens[i]: self.q_cu_seq_lens[i + 1]]
            table = self.block_tables[self.cu_blocks_lens[i]: self.cu_blocks_lens[i + 1]]
            out.append((q, kv, slots, table))
        return out


def golden_attention_cases():
    return sorted(GOLDEN.glob("attn_*.npz"))


@pytest.fixture(params=golden_attention_cases(), ids=lambda p: p.stem[len("attn_"):])
def golden_attention(request) -> GoldenAttention:
    return GoldenAttention(request.param)
