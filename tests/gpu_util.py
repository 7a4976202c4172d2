"""Helpers shared by the -m gpu parity tests (torch <-> numpy, bf16 views)."""
import numpy as np
import torch

from oracle import llama as LL

DEV = "cuda"


def to_bf16_t(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(DEV).to(torch.bfloat16)


def bf16_np(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()


def bits(t: torch.Tensor) -> np.ndarray:
    """raw bytes of a tensor as uint8 numpy (for bit-exact comparisons)."""
    return t.detach().contiguous().view(torch.uint8).cpu().numpy()


def rel_fro(a, b) -> float:
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def qtensor_np(q) -> np.ndarray:
    return q.data.detach().cpu().numpy()


def weights_to_oracle(w: dict) -> dict:
    """candle_vllm_b200.synthetic weights (torch, device) -> oracle.llama weights (numpy)."""
    qt = lambda q: (qtensor_np(q), q.ggml_type, q.shape[0], q.shape[1])
    out = dict(tok_embeddings=w["tok_embeddings"].cpu().numpy(), norm=w["norm"].cpu().numpy(), output=qt(w["output"]), layers=[])
    for lw in w["layers"]:
        out["layers"].append(dict(attn_norm=lw["attn_norm"].cpu().numpy(), ffn_norm=lw["ffn_norm"].cpu().numpy(),
                                  **{k: qt(lw[k]) for k in ("wq", "wk", "wv", "wo", "w1", "w2", "w3")}))
    return out
