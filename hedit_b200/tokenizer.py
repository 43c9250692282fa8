"""A dependency-free word-level tokenizer with the HF-CLIP call surface the P2P set-up code needs (`encode`,
`decode`, `__call__(...).input_ids`, `model_max_length`).  The real CLIP BPE vocabulary is an L0 third-party asset
that is not available offline; real deployments pass `model.tokenizer` instead.  Token ids are stable hashes."""
from __future__ import annotations

import zlib
from typing import List, Sequence, Union

import torch


class WordTokenizer:
    model_max_length = 77
    bos, eos = 49406, 49407

    def __init__(self):
        self._words = {}

    def _id(self, w: str) -> int:
        i = zlib.crc32(w.encode()) % 49000 + 1
        self._words.setdefault(i, w)
        return i

    def encode(self, text: str) -> List[int]:
        return [self.bos] + [self._id(w) for w in text.split(" ") if w] + [self.eos]

    def decode(self, ids: Sequence[int]) -> str:
        return " ".join("<|startoftext|>" if int(i) == self.bos else "<|endoftext|>" if int(i) == self.eos else self._words.get(int(i), "?")
                        for i in ids)

    def __call__(self, prompts: Union[str, List[str]], padding="max_length", max_length=77, truncation=True, return_tensors="pt"):
        prompts = [prompts] if isinstance(prompts, str) else prompts
        ids = torch.full((len(prompts), max_length), self.eos, dtype=torch.int64)
        for r, p in enumerate(prompts):
            e = self.encode(p)[:max_length]
            e[-1] = self.eos
            ids[r, : len(e)] = torch.tensor(e)
        return type("Enc", (), {"input_ids": ids})()
