"""Host-side mirror of the batched paths (include/zg_b200.h, zg_batch_*): B independent sequences forwarded as B
rows of every Linear -- the reference's GPT.forward / generate (main.zig:178-207, 322-342) per sequence, the
tensor-core GEMM and flash-attention kernels underneath.  Nothing here computes on the CPU."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import lib as _lib
from .gpt import GPT


class BatchEngine:
    def __init__(self, gpt: GPT, n_seqs: int, cache_rows: Optional[int] = None, max_prompt: int = 0, graph: bool = True,
                 tf32_single_pass: bool = False, general_gemm_only: bool = False, exact_prefill: bool = False,
                 storage16: bool = False):
        self.gpt, self.n_seqs = gpt, int(n_seqs)
        self.cache_rows = int(cache_rows or gpt.config.context_size)
        self.max_prompt = int(max_prompt)
        L = _lib.load()
        self._h = L.zg_batch_create(C.byref(gpt.c), self.n_seqs, self.cache_rows, self.max_prompt,
                                    (0 if graph else 1) | (2 if tf32_single_pass else 0) | (4 if exact_prefill else 0) |
                                    (8 if general_gemm_only else 0) | (16 if storage16 else 0))
        _lib.check()
        if not self._h:
            raise _lib.ZgError("zg_batch_create failed")
        self.pitch = int(L.zg_batch_logits_pitch(self._h))
        self.storage_bits = int(L.zg_batch_storage_bits(self._h))
        self.fused_argmax = bool(L.zg_batch_fused_argmax(self._h))  # greedy steps never write logits (stream-K path)

    def _tok(self, a, n) -> np.ndarray:
        t = np.ascontiguousarray(a, np.uint64).reshape(-1)
        assert t.size == n, (t.size, n)
        return t

    def forward(self, seq_len: int, tokens: Sequence[int], compute_logits=True) -> None:
        """compute_logits: False / True as in GPT.forward; 2 = next-token ids only (read_tokens), no logits written."""
        t = self._tok(tokens, self.n_seqs)
        _lib.load().zg_batch_forward(self._h, seq_len, t.ctypes.data_as(_lib.c_size_p), int(compute_logits))
        _lib.check()

    def prefill(self, tokens: np.ndarray, compute_logits: bool = True) -> None:
        tokens = np.asarray(tokens)
        assert tokens.ndim == 2 and tokens.shape[0] == self.n_seqs
        t = self._tok(tokens, tokens.size)
        rc = _lib.load().zg_batch_prefill(self._h, t.ctypes.data_as(_lib.c_size_p), tokens.shape[1], int(compute_logits))
        _lib.check()
        if rc:
            raise _lib.ZgError(f"zg_batch_prefill -> {rc}")

    def prefill_resident(self, T: int, compute_logits: bool = True) -> None:
        _lib.load().zg_batch_prefill_resident(self._h, T, int(compute_logits))
        _lib.check()

    def logits(self) -> np.ndarray:
        L = _lib.load()
        out = np.empty((self.n_seqs, self.pitch), np.float32)
        L.zg_download(out.ctypes.data, L.zg_batch_logits(self._h), out.nbytes)
        _lib.check()
        return out[:, : self.gpt.config.vocab_size]

    def kv(self, layer: int, rows: int):
        L = _lib.load()
        E = self.gpt.config.n_embed
        k = np.empty((self.n_seqs, self.cache_rows, E), np.float16 if self.storage_bits == 16 else np.float32)
        v = np.empty_like(k)
        L.zg_download(k.ctypes.data, L.zg_batch_k_cache(self._h, layer), k.nbytes)
        L.zg_download(v.ctypes.data, L.zg_batch_v_cache(self._h, layer), v.nbytes)
        _lib.check()
        return k[:, :rows], v[:, :rows]

    def generate_greedy(self, prompts: np.ndarray, n_total: int, use_prefill: bool = False) -> np.ndarray:
        prompts = np.asarray(prompts)
        assert prompts.ndim == 2 and prompts.shape[0] == self.n_seqs
        p = self._tok(prompts, prompts.size)
        out = np.zeros(self.n_seqs * n_total, np.uint64)
        rc = _lib.load().zg_batch_generate_greedy(self._h, p.ctypes.data_as(_lib.c_size_p), prompts.shape[1], n_total,
                                                  out.ctypes.data_as(_lib.c_size_p), int(use_prefill))
        _lib.check()
        if rc:
            raise _lib.ZgError(f"zg_batch_generate_greedy -> {rc}")
        return out.reshape(self.n_seqs, n_total).astype(np.int64)

    def generate_sample(self, prompts: np.ndarray, n_total: int, temp: float, seed: int, seq_base: int = 0,
                        use_prefill: bool = False) -> np.ndarray:
        """generate() with GPT.sample per sequence; sequence b draws u = zg_philox_uniform(seed, step, seq_base + b)."""
        prompts = np.asarray(prompts)
        assert prompts.ndim == 2 and prompts.shape[0] == self.n_seqs
        p = self._tok(prompts, prompts.size)
        out = np.zeros(self.n_seqs * n_total, np.uint64)
        rc = _lib.load().zg_batch_generate_sample(self._h, p.ctypes.data_as(_lib.c_size_p), prompts.shape[1], n_total, temp,
                                                  seed, seq_base, out.ctypes.data_as(_lib.c_size_p), int(use_prefill))
        _lib.check()
        if rc:
            raise _lib.ZgError(f"zg_batch_generate_sample -> {rc}")
        return out.reshape(self.n_seqs, n_total).astype(np.int64)

    def read_tokens(self) -> np.ndarray:
        """argmax token of every sequence's last sampling step."""
        out = np.zeros(self.n_seqs, np.uint64)
        rc = _lib.load().zg_batch_read_tokens(self._h, out.ctypes.data_as(_lib.c_size_p))
        _lib.check()
        if rc:
            raise _lib.ZgError(f"zg_batch_read_tokens -> {rc}")
        return out.astype(np.int64)

    def set_position(self, pos: int) -> None:
        _lib.load().zg_batch_set_position(self._h, pos)
        _lib.check()

    def run_steps(self, n: int) -> None:
        _lib.load().zg_batch_run_steps(self._h, n)
        _lib.check()

    def close(self) -> None:
        if self._h:
            _lib.load().zg_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
