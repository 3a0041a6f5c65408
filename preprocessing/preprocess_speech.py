# -*- coding: UTF-8 -*-
"""Drop-in for the reference's preprocessing/preprocess_speech.py: same flags (--seed --ssl_type --save_path
--wav_dir --num_workers --n_layer --use_average), same <basename>.pt [T, D] float32 outputs; the encoder
forward runs in libserenc (hand-written sm_100a CUDA) instead of HuggingFace transformers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from interspeech_ser_b200.cli import main_speech  # noqa: E402

if __name__ == "__main__":
    sys.exit(main_speech())
