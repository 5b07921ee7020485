"""token ids -> LaTeX text against vectors produced by the reference's RegExTokenizer / process_output."""
import json
import os

import numpy as np
import pytest

from texocr_b200.detok import Detokenizer, process_output

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def tk():
    with open(os.path.join(HERE, "golden", "golden_tokenizer_v1.json")) as f:
        return json.load(f)


def test_vocab_table_from_merges_equals_reference(tk):
    d = Detokenizer.from_merges(tk["bp_merges"], tk["special_tokens"])
    assert d.vocab_size == tk["vocab_size"] == 1000
    for k, v in tk["vocab_bytes"].items():
        assert d.vocab_bytes[int(k)] == bytes(v)


def test_decode_matches_reference(tk):
    d = Detokenizer.from_merges(tk["bp_merges"], tk["special_tokens"])
    for ids, text in zip(tk["random_ids"], tk["random_decoded"]):
        assert d.decode(ids) == text                       # includes split multi-byte characters -> U+FFFD
    for ids, text, proc, src in zip(tk["latex_ids"], tk["latex_decoded"], tk["latex_processed"], tk["latex"]):
        assert d.decode(ids) == text                       # the reference's own decode of its own encode
        if src.isascii():
            assert text == src                             # (non-ASCII characters split over byte tokens come back as U+FFFD)
        assert process_output(d.decode(ids)) == proc
    with pytest.raises(ValueError, match="not found"):
        d.decode([5, 4242])


def test_process_output_matches_reference(tk):
    for a, b in zip(tk["process_in"], tk["process_out"]):
        assert process_output(a) == b


def test_load_reference_file_format_and_batch_decode(tk, tmp_path):
    merges = {(a, b): t for a, b, t in tk["bp_merges"]}
    p = tmp_path / "tok.txt"
    p.write_text(f"{tk['vocab_size']}\n{tk['special_tokens']}\n{merges}\n")
    d = Detokenizer.load(str(p))
    eos, pad = tk["special_tokens"]["<EOS>"], tk["special_tokens"]["<PAD>"]
    T = max(len(r) for r in tk["latex_ids"]) + 3
    batch = np.full((len(tk["latex_ids"]), T), 17, dtype=np.int64)     # garbage after EOS, as generate produces
    for i, r in enumerate(tk["latex_ids"]):
        batch[i, :len(r)] = r
        batch[i, len(r)] = eos
    assert d.decode_batch(batch, eos_token=eos) == tk["latex_processed"]
    assert d.decode_batch(batch[:, :5], eos_token=eos, postprocess=False) == [d.decode(r[:5]) for r in tk["latex_ids"]]
    bad = tmp_path / "bad.txt"
    bad.write_text("1000\n__import__('os').system('true')\n{}\n")
    with pytest.raises(Exception):
        Detokenizer.load(str(bad))
