"""Text front-end (bit-exact ids) and collate contract -- host-side pieces either side of the hot path."""
import os

import numpy as np
import pytest
import torch


def test_readme_known_answer_vector(golden_dir):
    from text import text_to_sequence
    ref = np.load(os.path.join(golden_dir, "text_ids.npz"))["ids"].tolist()
    assert text_to_sequence("감정있는 한국어 목소리 생성", ["korean_cleaners"]) == ref      # README.md:16-24


def test_symbol_table_quirks():
    from text import sequence_to_text, text_to_sequence
    from text.symbols import kor_symbols
    assert len(kor_symbols) == 80 and kor_symbols.count("ㅇ") == 2
    ids = text_to_sequence("앙", ["korean_cleaners"])
    assert ids == [13, 21, 62, 1]                     # tail 'ㅇ' maps to the LAST occurrence (62), quirk Q5
    assert sequence_to_text(ids).endswith("~")
    assert text_to_sequence("사과 3개", ["korean_cleaners"]) == text_to_sequence("사과 세개", ["korean_cleaners"])


@pytest.mark.reference
def test_text_ids_match_reference_on_filelists():
    """every sentence of the reference's validation/test filelists through the reference's own text pipeline (run under
    the oracle shims) and through the restatement: identical id sequences"""
    import importlib
    import sys
    from oracle import ref_shims
    ref_shims._install_shims()
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "text" or k.startswith("text.") or k == "hparams"}
    sys.path.insert(0, "/root/reference")
    try:
        ref_fn = importlib.import_module("text").text_to_sequence
    finally:
        sys.path.remove("/root/reference")
        for k in list(sys.modules):
            if k == "text" or k.startswith("text.") or k == "hparams":
                sys.modules.pop(k)
        sys.modules.update(saved)
    from text import text_to_sequence
    n = 0
    for name in ("koemo_spk_emo_all_valid.txt", "koemo_spk_emo_all_test.txt"):
        for line in open(os.path.join("/root/reference/filelists", name), encoding="utf-8"):
            t = line.strip().split("|")[1]
            assert text_to_sequence(t, ["korean_cleaners"]) == ref_fn(t, ["korean_cleaners"]), t
            n += 1
    assert n > 1000


def test_collate_contract():
    from data_utils import TextMelCollate
    g = torch.Generator().manual_seed(0)
    items = []
    for n_txt, n_mel in ((5, 17), (9, 30), (7, 22)):
        items.append((torch.randint(2, 79, (n_txt,), generator=g).int(), torch.randn(80, n_mel, generator=g),
                      torch.tensor([1.0]), torch.tensor([0.0, 0.0, 1.0, 0.0])))
    text, in_len, mel, gate, out_len, spk, emo = TextMelCollate(1)(items)
    assert in_len.tolist() == [9, 7, 5] and out_len.tolist() == [30, 22, 17]
    assert text.shape == (3, 9) and mel.shape == (3, 80, 30) and gate.shape == (3, 30)
    assert float(text[2, 5:].sum()) == 0 and float(mel[2, :, 17:].abs().sum()) == 0
    assert gate[1].tolist() == [0.0] * 21 + [1.0] * 9                       # 1 from the last real frame on
    assert emo.dtype == torch.long and emo[0].tolist() == [0, 0, 1, 0]
