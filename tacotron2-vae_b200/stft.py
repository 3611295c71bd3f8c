"""reference stft.py -> the t2v STFT (forward transform only; the inverse/Griffin-Lim path is out of scope)."""
from t2v.frontend import STFT  # noqa: F401
