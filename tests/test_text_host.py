"""CPU tests of the text branch (preprocessing/preprocess_roberta.py): byte-level BPE tokenizer against HF's
RobertaTokenizer on the same synthetic vocabulary, the RoBERTa oracle against HF golden vectors, weight-name mapping,
and the Whisper LoRA checkpoint layout of preprocess_whisper_pretrained.py."""
import json
import os

import numpy as np
import pytest
import torch

from interspeech_ser_b200 import configs
from interspeech_ser_b200.text import RobertaTokenizer, bytes_to_unicode, _valid_lengths
from interspeech_ser_b200.weights import from_hf_state_dict, random_init
from oracle import ssl_oracle as O

CORPUS = [
    "I can't believe it's already over!",
    "well, that was   unexpected... wasn't it?",
    "The quick brown fox jumps over the lazy dog 1234 times.",
    "naïve café — déjà vu; 你好 🙂",
    "",
    " leading and trailing spaces  ",
    "so so so so so happy happy happy, I'm I'm I'm",
    "um, uh, I - I don't know. maybe? maybe not!",
]


def train_tiny_bpe(texts, n_merges=120):
    """A small byte-level BPE (GPT-2 style) trained on `texts`: (vocab dict, merges list)."""
    import regex
    from collections import Counter

    pat = regex.compile(r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""")
    b2u = bytes_to_unicode()
    words = Counter()
    for t in texts:
        for tok in pat.findall(t):
            words[tuple(b2u[b] for b in tok.encode("utf-8"))] += 1
    merges = []
    for _ in range(n_merges):
        pairs = Counter()
        for w, c in words.items():
            for a, b in zip(w, w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        (a, b), _ = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
        merges.append(f"{a} {b}")
        new = Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i < len(w) - 1 and w[i] == a and w[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            new[tuple(out)] += c
        words = new
    vocab = {"<s>": 0, "<pad>": 1, "</s>": 2, "<unk>": 3}
    for ch in b2u.values():
        vocab.setdefault(ch, len(vocab))
    for m in merges:
        vocab.setdefault(m.replace(" ", ""), len(vocab))
    vocab["<mask>"] = len(vocab)
    return vocab, merges


@pytest.fixture(scope="module")
def tok_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("tok")
    vocab, merges = train_tiny_bpe(CORPUS * 3)
    with open(d / "vocab.json", "w", encoding="utf-8") as fh:
        json.dump(vocab, fh, ensure_ascii=False)
    with open(d / "merges.txt", "w", encoding="utf-8") as fh:
        fh.write("#version: 0.2\n" + "\n".join(merges) + "\n")
    return str(d)


def test_tokenizer_matches_hf_roberta_tokenizer(tok_files):
    """Same ids and masks as transformers' RobertaTokenizer (the class preprocess_roberta.py:103 loads) with the call
    the script makes (:49-55): padding='max_length', truncation=True, max_length=MAX_LEN, return_tensors='pt'."""
    tr = pytest.importorskip("transformers")
    hf = tr.RobertaTokenizer(vocab=os.path.join(tok_files, "vocab.json"), merges=os.path.join(tok_files, "merges.txt"))
    mine = RobertaTokenizer.from_pretrained(tok_files)
    assert (mine.bos_token_id, mine.pad_token_id, mine.eos_token_id) == (0, 1, 2)
    extra = ["unseen words like xylophone & zebra?", "x" * 300, "tabs\tand\nnewlines"]
    for max_len in (80, 8):
        for t in CORPUS + extra:
            a = hf(t, padding="max_length", truncation=True, max_length=max_len, return_tensors="pt")
            b = mine(t, padding="max_length", truncation=True, max_length=max_len, return_tensors="pt")
            assert b["input_ids"].shape == (1, max_len)
            assert torch.equal(a["input_ids"], b["input_ids"]), t
            assert torch.equal(a["attention_mask"], b["attention_mask"]), t
    batch = mine(CORPUS, padding="max_length", truncation=True, max_length=16, return_tensors="pt")
    ref = hf(CORPUS, padding="max_length", truncation=True, max_length=16, return_tensors="pt")
    assert torch.equal(batch["input_ids"], ref["input_ids"]) and torch.equal(batch.attention_mask, ref["attention_mask"])
    with pytest.raises(OSError):
        RobertaTokenizer.from_pretrained("no/such-tokenizer")


def test_valid_lengths_accepts_only_right_padded_batches():
    ids = torch.tensor([[0, 5, 6, 2, 1, 1], [0, 7, 2, 1, 1, 1]])
    assert _valid_lengths(ids, ids.ne(1).long(), 1) == [4, 3]
    with pytest.raises(NotImplementedError):
        _valid_lengths(torch.tensor([[0, 1, 6, 2, 1, 1]]), None, 1)          # pad inside the sequence
    with pytest.raises(NotImplementedError):
        _valid_lengths(ids, torch.ones_like(ids), 1)                          # mask disagrees with the pad tokens
    assert _valid_lengths(torch.tensor([[0, 5, 6, 2]]), None, 1) == [4]


def test_roberta_oracle_matches_hf_golden(golden_dir):
    from oracle.make_golden import synth_token_rows

    path = os.path.join(golden_dir, "tiny__roberta.npz")
    g = np.load(path)
    cfg = configs.get_config("tiny/roberta")
    w = random_init(cfg, int(g["seed"]))
    lengths = [int(n) for n in g["lengths"]]
    T = int(g["max_len"])
    rows = synth_token_rows(cfg, int(g["ids_seed"]), lengths, T)
    for j, n in enumerate(lengths):
        hs = O.roberta_hidden_states(cfg, w, rows[j])
        assert len(hs) == cfg.num_hidden_layers + 1 and hs[0].shape == (T, cfg.hidden_size)
        np.testing.assert_allclose(np.stack([h[:n].mean(0).numpy() for h in hs]), g[f"pooled_{j}"], atol=2e-5, rtol=1e-4)
        np.testing.assert_allclose(np.stack([h.mean(0).numpy() for h in hs]), g[f"pooled_all_{j}"], atol=2e-5, rtol=1e-4)
        np.testing.assert_allclose(hs[-1][[0, 1, n - 1, T - 1]].numpy(), g[f"last_{j}"], atol=5e-5, rtol=1e-4)


def test_roberta_state_dict_conversion_roundtrip():
    tr = pytest.importorskip("transformers")
    from oracle.make_golden import hf_roberta

    cfg = configs.get_config("tiny/roberta")
    w = random_init(cfg, 4)
    m = hf_roberta(cfg, w)                   # asserts from_hf_state_dict(m.state_dict()) == w
    wrapped = {"roberta." + k: v for k, v in m.state_dict().items()}     # e.g. RobertaForSequenceClassification checkpoints
    back = from_hf_state_dict(cfg, wrapped)
    assert set(back) == set(w)
    for k in w:
        assert np.array_equal(back[k], w[k]), k
    assert isinstance(m, tr.RobertaModel)


def test_whisper_lora_checkpoint_layout_is_merged():
    """preprocess_whisper_pretrained.py:115-138,180-181: WhisperAudioClassifier state dict = `whisper.base_model.model.*`
    (peft r=8 / alpha=16 on every q_proj and v_proj, decoder included) + `classifier.*`. The converter must strip the
    prefix, merge the encoder adapters into the k-bias-less Whisper layout and ignore decoder / classifier keys."""
    pytest.importorskip("transformers")
    from oracle.make_golden import hf_model

    cfg = configs.get_config("tiny/whisper")
    w = random_init(cfg, 2)
    enc = hf_model(cfg, w)
    rng = np.random.default_rng(9)
    sd, merged = {}, {k: v.copy() for k, v in w.items()}
    for k, v in enc.state_dict().items():
        mod, _, leaf = k.rpartition(".")
        pre = "whisper.base_model.model.encoder."
        sd[pre + (f"{mod}.base_layer.{leaf}" if mod.endswith(("q_proj", "v_proj")) else k)] = v
    d = cfg.hidden_size
    for i in range(cfg.num_hidden_layers):
        for s_, name in (("q", "q_proj"), ("v", "v_proj")):
            a = (rng.standard_normal((8, d)) * 0.2).astype(np.float32)
            b = (rng.standard_normal((d, 8)) * 0.2).astype(np.float32)
            mod = f"whisper.base_model.model.encoder.layers.{i}.self_attn.{name}"
            sd[mod + ".lora_A.default.weight"], sd[mod + ".lora_B.default.weight"] = torch.from_numpy(a), torch.from_numpy(b)
            merged[f"layer{i}.{s_}.weight"] = (w[f"layer{i}.{s_}.weight"].astype(np.float64) + 2.0 * (b.astype(np.float64) @ a.astype(np.float64))).astype(np.float32)
    # decoder adapters and the classifier head are present in the real checkpoint and must be ignored
    sd["whisper.base_model.model.decoder.layers.0.self_attn.q_proj.base_layer.weight"] = torch.zeros(d, d)
    sd["whisper.base_model.model.decoder.layers.0.self_attn.q_proj.lora_A.default.weight"] = torch.zeros(8, d)
    sd["whisper.base_model.model.decoder.layers.0.self_attn.q_proj.lora_B.default.weight"] = torch.zeros(d, 8)
    sd["classifier.0.weight"] = torch.zeros(512, d)
    got = from_hf_state_dict(cfg, sd)
    assert set(got) == set(w)
    assert not any(k.endswith("k.bias") for k in got)
    for k in w:
        np.testing.assert_allclose(got[k], merged[k], atol=1e-6, err_msg=k)
    assert float(np.abs(got["layer0.q.weight"] - w["layer0.q.weight"]).max()) > 0.1     # the adapter matters
