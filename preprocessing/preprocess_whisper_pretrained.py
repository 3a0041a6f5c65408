# -*- coding: UTF-8 -*-
"""Drop-in for the reference's preprocessing/preprocess_whisper_pretrained.py: embedding extraction with a LoRA-tuned
Whisper-large-v3 (peft r=8, alpha=16 on q_proj/v_proj, :115-138; state dict under `whisper.base_model.model.*`,
:180-181).  The reference hard-codes the checkpoint path (:180); here it is --checkpoint.  The adapters are folded into
the dense q/v weights at load (weights.merge_lora; decoder and classifier keys are dropped), so the same sm_100a kernels
run with no extra GEMMs.  Everything else is preprocess_whisper.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from interspeech_ser_b200.cli import main_whisper  # noqa: E402

if __name__ == "__main__":
    if not any(a == "--checkpoint" or a.startswith("--checkpoint=") for a in sys.argv[1:]):
        print("Error: --checkpoint <lora state dict .pt> is required")
        sys.exit(1)
    sys.exit(main_whisper())
