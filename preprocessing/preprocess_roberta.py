# -*- coding: UTF-8 -*-
"""Drop-in for the reference's preprocessing/preprocess_roberta.py: same flags (--roberta_type --df_path --save_path
--num_workers --max_len --use_average), same <basename>.pt outputs ([max_len, D] fp32: last_hidden_state, or the
mean of the last four hidden states with --use_average y). Tokenisation is byte-level BPE on the host; embeddings and
the 24 post-LN layers run in libserenc (hand-written sm_100a CUDA)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from interspeech_ser_b200.cli import main_roberta  # noqa: E402

if __name__ == "__main__":
    sys.exit(main_roberta())
