# -*- coding: UTF-8 -*-
"""Drop-in for the reference's preprocessing/preprocess_whisper.py: same flags, same <basename>.pt outputs
(first min(ceil(len/320), hidden_size) encoder frames of the selected hidden state); log-mel frontend and encoder
run in libserenc (hand-written sm_100a CUDA)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from interspeech_ser_b200.cli import main_whisper  # noqa: E402

if __name__ == "__main__":
    sys.exit(main_whisper())
