"""Symbol inventories (the vocabulary contract of the checkpoints; reference text/symbols.py:21, text/korean.py:24)."""
_pad, _end = "_", "~"
_punct = "!'(),-.:;? "

# 80 symbols of the Korean model: pad, eos, 19 lead consonants (Hangul Jamo block), 21 vowels and 30 tail consonants
# (Hangul Compatibility Jamo block -- the tail 'ㅇ' therefore appears twice, see quirk Q5), punctuation + space.
_LEADS = "".join(chr(c) for c in range(0x1100, 0x1113))
_VOWELS_HCJ = "".join(chr(c) for c in range(0x314F, 0x3164))
_TAILS_HCJ = "ㄱㄲㄳㄴㄵㄶㅇㄹㄺㄻㄼㄽㄾㄿㅀㅁㅂㅄㅅㅆㅇㅈㅊㅋㅌㅍㅎ"
kor_symbols = _pad + _end + _LEADS + _VOWELS_HCJ + _TAILS_HCJ + _punct
assert len(kor_symbols) == 80

_letters = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"
eng_symbols = [_pad] + list("-") + list("!'(),.:;? ") + list(_letters) + [_end]
