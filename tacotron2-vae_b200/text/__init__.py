"""text -> id sequence (reference text/__init__.py:30-60).  Bit-exact contract: the README known-answer vector."""
import re

from .korean import tokenize as _ko_tokenize
from .symbols import eng_symbols, kor_symbols

_curly_re = re.compile(r"(.*?)\{(.+?)\}(.*)")


def _tables(cleaner_names):
    symbols = kor_symbols if list(cleaner_names) == ["korean_cleaners"] else eng_symbols
    # dict comprehension: a symbol that occurs twice keeps its LAST index (the tail 'ㅇ' -> 62, quirk Q5)
    return {s: i for i, s in enumerate(symbols)}, {i: s for i, s in enumerate(symbols)}


def _clean(text, cleaner_names):
    for name in cleaner_names:
        if name == "korean_cleaners":
            text = _ko_tokenize(text)
        elif name in ("basic_cleaners", "transliteration_cleaners", "english_cleaners"):
            text = re.sub(r"\s+", " ", text.lower())
        else:
            raise Exception("Unknown cleaner: %s" % name)
    return text


def text_to_sequence(text, cleaner_names):
    to_id, _ = _tables(cleaner_names)
    keep = lambda s: s in to_id and s != "_" and s != "~"
    seq = []
    while text:
        m = _curly_re.match(text) if isinstance(text, str) else None
        if not m:
            seq += [to_id[s] for s in _clean(text, cleaner_names) if keep(s)]
            break
        seq += [to_id[s] for s in _clean(m.group(1), cleaner_names) if keep(s)]
        seq += [to_id[s] for s in ("@" + a for a in m.group(2).split()) if keep(s)]
        text = m.group(3)
    seq.append(to_id["~"])
    return seq


def sequence_to_text(sequence, cleaner_names=("korean_cleaners",)):
    _, to_sym = _tables(list(cleaner_names))
    return "".join(to_sym.get(i, "") for i in sequence)
