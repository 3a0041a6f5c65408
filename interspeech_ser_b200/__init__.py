"""B200-native speech-SSL embedding extraction (the hot path of AI-Unicamp/interspeech_ser's preprocessing scripts).

Public surface (mirrors what the reference imports from `transformers`):

    from interspeech_ser_b200 import AutoModel, AutoFeatureExtractor, AutoProcessor

The arithmetic lives in libserenc.so (csrc/, hand-written sm_100a CUDA behind the C ABI of include/serenc.h).
Importing this package does not need a GPU; creating a model does, and fails loudly without one.
"""
from .configs import EncoderConfig, get_config  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name in ("AutoModel", "AutoFeatureExtractor", "AutoProcessor", "SpeechEncoderModel", "WhisperModel"):
        from . import modeling
        return getattr(modeling, name)
    if name in ("RobertaModel", "RobertaTokenizer"):   # text branch (preprocessing/preprocess_roberta.py)
        from . import text
        return getattr(text, name)
    raise AttributeError(name)
