"""Korean text -> jamo tokens (restatement of reference text/korean.py:177-394 without the `jamo`/`nltk` packages).

Pipeline: normalise (quotes, bracketed annotations, upper-case letter names, numbers -> Korean reading) ->
decompose precomposed syllables U+AC00..U+D7A3 by Unicode arithmetic (what jamo.hangul_to_jamo does) -> map every
jamo to the model's symbol (leads stay in the Hangul Jamo block, vowels/tails move to the compatibility block).
The substitution dictionaries of the reference (ko_dictionary.py: English loan words, units, abbreviations) are data,
not algorithm, and are not reproduced: such words pass through unchanged ("parity unpinned" beyond the README vector)."""
import ast
import re

from .symbols import _LEADS, _TAILS_HCJ, _VOWELS_HCJ

EOS = "~"
_VOWELS_J = "".join(chr(c) for c in range(0x1161, 0x1176))
_TAILS_J = "".join(chr(c) for c in range(0x11A8, 0x11C3))
# tail jamo U+11A8.. in Unicode order -> compatibility letters; the reference's table (korean.py:150-151) sends the tail
# digraph 'ᆮ' (U+11AE) to 'ㅇ' as well, which is reproduced here
_TAIL_MAP = dict(zip(_TAILS_J, _TAILS_HCJ))
_VOWEL_MAP = dict(zip(_VOWELS_J, _VOWELS_HCJ))
_HCJ_LEADS = "ㄱㄲㄴㄷㄸㄹㅁㅂㅃㅅㅆㅇㅈㅉㅊㅋㅌㅍㅎ"
_HCJ_TO_LEAD = dict(zip(_HCJ_LEADS, _LEADS))
_UPPER = dict(zip("ABCDEFGHIJKLMNOPQRSTUVWXYZ",
                  "에이 비 씨 디 이 에프 지 에이치 아이 제이 케이 엘 엠 엔 오 피 큐 알 에스 티 유 브이 더블유 엑스 와이 지".split()))
_NUM = [""] + list("일이삼사오육칠팔구")
_UNIT4 = [""] + list("만억조경해")
_UNIT = [""] + list("십백천")
_COUNT = ["", "한", "두", "세", "네", "다섯", "여섯", "일곱", "여덟", "아홉"]
_COUNT_TENS = {"십": "열", "두십": "스물", "세십": "서른", "네십": "마흔", "다섯십": "쉰", "여섯십": "예순", "일곱십": "일흔",
               "여덟십": "여든", "아홉십": "아흔"}
_COUNTERS = "(시|명|가지|살|마리|포기|송이|수|톨|통|점|개|벌|척|채|다발|그루|자루|줄|켤레|그릇|잔|마디|상자|사람|곡|병|판)"
_NUMBER = r"([+-]?\d[\d,]*)[\.]?\d*"


def decompose(text):
    """precomposed Hangul syllables -> lead/vowel/tail jamo (Unicode 3.0 arithmetic)"""
    out = []
    for ch in text:
        o = ord(ch) - 0xAC00
        if 0 <= o < 11172:
            out.append(chr(0x1100 + o // 588))
            out.append(chr(0x1161 + (o % 588) // 28))
            if o % 28:
                out.append(chr(0x11A7 + o % 28))
        elif ch in _HCJ_TO_LEAD:                     # a bare compatibility consonant is read as a lead (korean.py:183)
            out.append(_HCJ_TO_LEAD[ch])
        else:
            out.append(ch)
    return out


def read_number(m, is_count=False):
    num_str, unit = (m.group(1), m.group(2)) if is_count else (m.group(), "")
    num_str = num_str.replace(",", "")
    try:
        num = ast.literal_eval(num_str)
    except Exception:   # noqa: BLE001
        num = int(num_str)
    if num == 0:
        return "영"
    parts = num_str.split(".")
    digits, frac = (parts[0], parts[1]) if len(parts) == 2 else (parts[0], None)
    digits = str(abs(int(digits)))
    kor, tmp, size = "", [], len(digits)
    for i, v in enumerate(digits, start=1):
        v = int(v)
        if v:
            tmp += (_COUNT if is_count else _NUM)[v]
            tmp += _UNIT[(size - i) % 4]
        if (size - i) % 4 == 0 and tmp:
            kor += "".join(tmp) + _UNIT4[(size - i) // 4]
            tmp = []
    if is_count:
        if kor.startswith("한") and len(kor) > 1:
            kor = kor[1:]
        kor = re.sub("|".join(_COUNT_TENS), lambda x: _COUNT_TENS[x.group()], kor)
    elif kor.startswith("일") and len(kor) > 1:
        kor = kor[1:]
    if frac is not None:
        kor += "쩜 " + "".join(("영" + "".join(_NUM[1:]))[int(c)] for c in frac)
    if num_str.startswith("+"):
        kor = "플러스 " + kor
    elif num_str.startswith("-"):
        kor = "마이너스 " + kor
    return kor + unit


def normalize(text):
    text = text.strip().replace("'", "").replace('"', "")
    text = re.sub(r"\(\d+일\)", "", text)
    text = re.sub(r"[A-Z]+", lambda m: "".join(_UPPER[c] for c in m.group()), text)
    text = re.sub(_NUMBER + _COUNTERS, lambda m: read_number(m, True), text)
    text = re.sub(_NUMBER, lambda m: read_number(m, False), text)
    return text


def tokenize(text):
    """-> list of model symbols + EOS (symbol_type 1 of the reference)"""
    toks = []
    for j in decompose(normalize(text)):
        toks.append(_VOWEL_MAP.get(j) or _TAIL_MAP.get(j) or j)
    return toks + [EOS]
