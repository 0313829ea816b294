"""GPTConfig (reference: src/main.zig:5-23) and the GPT-2 size table.

The reference hard-codes 124M (main.zig:346); the other sizes are the published GPT-2
family with head_dim 64, vocab 50257 and context 1024 throughout.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class GPTConfig:
    vocab_size: int = 50257
    context_size: int = 1024
    n_layer: int = 12
    n_heads: int = 12
    n_embed: int = 768

    @property
    def head_dim(self) -> int:
        return self.n_embed // self.n_heads

    def n_params(self) -> int:
        E, L, V, C = self.n_embed, self.n_layer, self.vocab_size, self.context_size
        return V * E + C * E + L * (12 * E * E + 13 * E) + 2 * E

    def weight_bytes_per_token(self, elem: int = 4) -> int:
        """Algorithmic weight bytes one decode step must read: all block
        weights + ln_f + the tied lm_head + the two gathered embedding rows."""
        E, L, V = self.n_embed, self.n_layer, self.vocab_size
        return elem * (L * (12 * E * E + 13 * E) + 2 * E + V * E + 2 * E)

    def kv_bytes_per_token(self, seq_len: int, batch: int = 1, elem: int = 4) -> int:
        """Read seq_len-1 cached rows and write 1 new row of K and V per layer per sequence."""
        return elem * batch * self.n_layer * 2 * self.n_embed * seq_len

    def decode_bytes(self, seq_len: int, batch: int = 1, elem: int = 4, fused_argmax: bool = True) -> int:
        tail = 8 * batch if fused_argmax else 4 * batch * self.vocab_size
        return self.weight_bytes_per_token(elem) + self.kv_bytes_per_token(seq_len, batch, elem) + tail

    def prefill_flops(self, batch: int, T: int) -> float:
        E, L, V, H = self.n_embed, self.n_layer, self.vocab_size, self.n_heads
        return batch * T * 24.0 * E * E * L + batch * L * H * 4.0 * self.head_dim * T * (T + 1) / 2 + batch * 2.0 * V * E


SIZES = {
    "124M": GPTConfig(50257, 1024, 12, 12, 768),
    "355M": GPTConfig(50257, 1024, 24, 16, 1024),
    "774M": GPTConfig(50257, 1024, 36, 20, 1280),
    "1.5B": GPTConfig(50257, 1024, 48, 25, 1600),
}
SIZE_INDEX = {"124M": 0, "355M": 1, "774M": 2, "1.5B": 3}
