"""Token ids -> LaTeX text: the step right after the generate loop.

Host-side mirror of the reference's ``RegExTokenizer.decode`` (tokenizer/tokenizer.py:222-238), the wrapper's EOS
stripping (model/ocr_model.py:103-105) and ``process_output`` (utils.py:73-79).  Pure Python / regex like the
reference; a 512 x 256 id matrix decodes in a few tens of milliseconds, off the GPU's critical path.
"""
import ast
import re
from typing import Dict, Iterable, List, Optional, Sequence


def process_output(output: str) -> str:
    """utils.py:73-79: drop the tokeniser's whitespace, keeping one space after a control word that is followed by a
    letter or digit (``\\alpha x`` must not become ``\\alphax``)."""
    output = re.sub(r"(\\[a-zA-Z]+)\s+([a-zA-Z0-9])", r"\1<SPACE>\2", output)
    output = re.sub(r"\s+", "", output)
    return output.replace("<SPACE>", " ")


class Detokenizer:
    """id -> text table of a trained byte-pair tokenizer (tokenizer/tokenizer.py:11-33 for the table construction)."""

    def __init__(self, vocab_bytes: Dict[int, bytes], special_tokens: Optional[Dict[str, int]] = None,
                 vocab_size: Optional[int] = None):
        self.special_tokens = dict(special_tokens or {})
        self.vocab_bytes = {int(k): bytes(v) for k, v in vocab_bytes.items()}
        for tok, tid in self.special_tokens.items():
            self.vocab_bytes[int(tid)] = tok.encode("utf-8")
        # the reference decodes every token on its own (errors='replace'), so a multi-byte character split over two
        # tokens becomes replacement characters: keep that behaviour by precomputing per-token strings
        self.pieces = {k: v.decode("utf-8", errors="replace") for k, v in self.vocab_bytes.items()}
        # the model is built for the DECLARED vocabulary (line 1 of the tokenizer file, model/ocr_model.py:76): byte-pair
        # training may stop early (tokenizer/tokenizer.py `if not stats: break`), so the table can hold fewer entries while the
        # special ids still sit at vocab_size - 1 and below
        self.vocab_size = len(self.vocab_bytes) if vocab_size is None else int(vocab_size)
        if self.vocab_bytes and max(self.vocab_bytes) >= self.vocab_size:
            raise ValueError(f"token id {max(self.vocab_bytes)} does not fit a vocabulary of {self.vocab_size}")

    @classmethod
    def from_merges(cls, merges: Iterable[Sequence[int]], special_tokens: Optional[Dict[str, int]] = None,
                    vocab_size: Optional[int] = None) -> "Detokenizer":
        """merges: (left id, right id, new id) in training order; ids 0..255 are the raw bytes."""
        vocab = {i: bytes([i]) for i in range(256)}
        for left, right, new in merges:
            vocab[int(new)] = vocab[int(left)] + vocab[int(right)]
        return cls(vocab, special_tokens, vocab_size)

    @classmethod
    def load(cls, path: str) -> "Detokenizer":
        """Reads the reference's tokenizer file (tokenizer/tokenizer.py:110-126): three lines -- vocabulary size, the
        special-token dict, the merge dict {(left, right): new_id} -- parsed as literals (never evaluated)."""
        with open(path, "r") as f:
            vocab_size = int(f.readline())
            special = ast.literal_eval(f.readline().strip())
            merges = ast.literal_eval(f.readline().strip())
        if not isinstance(special, dict) or not isinstance(merges, dict):
            raise ValueError(f"{path}: not a tokenizer file (expected two dict literals after the vocabulary size)")
        if 256 + len(merges) + len(special) > vocab_size:
            raise ValueError(f"{path}: {256 + len(merges) + len(special)} entries for a declared vocabulary of {vocab_size}")
        return cls.from_merges([(l, r, t) for (l, r), t in merges.items()], special, vocab_size)

    def decode(self, tokens: Iterable[int]) -> str:
        """RegExTokenizer.decode: concatenation of the per-token strings; unknown ids raise like the reference."""
        try:
            return "".join(self.pieces[int(t)] for t in tokens)
        except KeyError as e:
            raise ValueError(f"Token {e.args[0]} not found in vocabulary.") from None

    def decode_batch(self, ids, eos_token: Optional[int] = None, postprocess: bool = True) -> List[str]:
        """ids: (B, T) tensor / array / nested list from ``generate``.  Every row is cut before its first ``eos_token``
        (rows keep generating until the slowest row has finished, model/decoder.py:115-118), decoded, and passed
        through ``process_output`` like TeXOCRWrapper.__call__ (model/ocr_model.py:100-110)."""
        rows = ids.tolist() if hasattr(ids, "tolist") else ids
        out = []
        for row in rows:
            if eos_token is not None and eos_token in row:
                row = row[: row.index(eos_token)]
            text = self.decode(row)
            out.append(process_output(text) if postprocess else text)
        return out
