"""Generate tests/golden/* by running the REFERENCE's own code in the build container.

    PYTHONPATH=/root/reference python scripts/make_golden.py

What runs verbatim from /root/reference (it cannot travel to the GPU box, so the outputs are committed):
  * ultravox.model.ultravox_processing.UltravoxProcessor.__call__ / DataCollatorForSeq2SeqWithAudio  (integer path)
    with transformers.WhisperFeatureExtractor and a deterministic stub tokenizer (the LFS Llama-3 tokenizer asset is
    a pointer file here);
  * ultravox.model.ultravox_model.{UltravoxProjector, StackAudioFrames, SwiGLU, RMSNorm} and the splice loop of
    UltravoxModel._prepare_audio_embeds  (float path; `accelerate` / `peft` are absent, so two empty stand-in modules
    are put in sys.modules AFTER importing transformers - none of the executed code touches them).
  * transformers' WhisperFeatureExtractor (third-party arithmetic of the path) on seeded waveforms -> log-mel goldens.
  * ultravox.model.ultravox_data_proc.UltravoxDataproc._process / _compute_loss_mask_len  (host feed) with the
    reference's real VoiceSample (ultravox.data.data_sample; `librosa` / `soundfile` are only touched by its file loaders,
    so empty stand-in modules suffice).  The package `ultravox.data` itself needs simple_parsing / datasets / streaming, so
    a stand-in module exposing the real VoiceSample and a restated 10-line `Dataproc` base is registered under that name.
"""
import json
import os
import sys
import types

import numpy as np
import torch
import transformers

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
sys.path.insert(0, REF)
os.makedirs(OUT, exist_ok=True)

for name in ("accelerate", "peft"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)

from ultravox.model import ultravox_config, ultravox_processing  # noqa: E402
from ultravox.model import ultravox_model  # noqa: E402


class StubTokenizer:
    """Deterministic word-hash tokenizer: enough for the processor's placeholder bookkeeping."""
    eos_token = "<|eot_id|>"
    eos_token_id = 128009
    pad_token_id = None
    padding_side = "right"
    model_input_names = ["input_ids", "attention_mask"]

    def get_vocab(self):
        return {self.eos_token: self.eos_token_id}

    def __call__(self, parts, add_special_tokens=False, **kw):
        return {"input_ids": [[(sum(map(ord, w)) * 31 + len(w)) % 100000 for w in p.split()] for p in parts]}

    def pad(self, features, padding=True, max_length=None, pad_to_multiple_of=None, return_tensors=None, **kw):
        L = max(len(f["input_ids"]) for f in features)
        out = {"input_ids": [], "attention_mask": []}
        for f in features:
            ids = list(map(int, f["input_ids"]))
            n = L - len(ids)
            if self.padding_side == "left":
                out["input_ids"].append([self.pad_token_id] * n + ids)
                out["attention_mask"].append([0] * n + [1] * len(ids))
            else:
                out["input_ids"].append(ids + [self.pad_token_id] * n)
                out["attention_mask"].append([1] * len(ids) + [0] * n)
        extra = {k: [f[k] for f in features] for k in features[0] if k not in ("input_ids", "attention_mask")}
        out.update(extra)
        return transformers.BatchFeature(out, tensor_type=return_tensors)


def make_processor(n_mels):
    fe = transformers.WhisperFeatureExtractor(feature_size=n_mels)

    class AP:
        feature_extractor = fe
        model_input_names = ["input_features"]

        def __call__(self, *a, **k):
            return fe(*a, **k)

    P = ultravox_processing.UltravoxProcessor
    p = P.__new__(P)
    tok = StubTokenizer()
    p.audio_padding, p.encoder_ds_factor, p.stack_factor = "longest", 2, 8
    p.audio_placeholder, p.audio_context_size = "<|audio|>", 3000
    p.tokenizer, p.vocab, p.audio_token_replacement = tok, tok.get_vocab(), tok.eos_token
    tok.pad_token_id = tok.eos_token_id
    p.audio_processor = AP()
    return p


def wave(i, n):
    return np.random.default_rng(1000 + i).standard_normal(n).astype(np.float32)


def tolist(bf, skip=("audio_values",)):
    return {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in bf.items() if k not in skip}


def processor_cases():
    p = make_processor(80)
    sr = 16000
    cases = []
    specs = [
        ("text_only", "Hello how are you", []),
        ("one_1s", "Test with <|audio|>", [sr]),
        ("one_10s", "Test with <|audio|> tail words", [10 * sr]),
        ("one_30s", "a b c d e f g h <|audio|> i j k l m", [30 * sr]),
        ("one_35s", "Test with <|audio|>", [35 * sr]),
        ("one_60s", "x <|audio|>", [60 * sr]),
        ("two", "Test with <|audio|> and <|audio|>", [sr, 10 * sr]),
        ("three_overflow", "Test with <|audio|> and <|audio|> and <|audio|>", [sr, 35 * sr, 10 * sr]),
        ("ragged", "<|audio|> mid <|audio|>", [16000 + 77, 3 * sr - 1]),
    ] + [(f"short_{n}", "<|audio|>", [n]) for n in (0, 1, 159, 160, 161, 319, 320, 321)]
    for name, text, lens in specs:
        audios = [wave(i, n) for i, n in enumerate(lens)]
        kw = dict(audios=audios, sampling_rate=sr, include_audio_num_chunks=True) if audios else {}
        r = p(text, **kw)
        d = tolist(r)
        if "audio_values" in r:
            d["audio_values_shape"] = list(r["audio_values"].shape)
        cases.append({"name": name, "text": text, "sample_counts": lens, "out": d})
    errors = []
    for text, lens in [("Hello <|audio|>", []), ("Hello <|audio|><|audio|>", [sr]), ("Hello", [sr]),
                       ("Hello <|audio|>", [sr, sr]), ("Hello <|audio|><|audio|>", [35 * sr])]:
        try:
            p(text, audios=[wave(i, n) for i, n in enumerate(lens)], sampling_rate=sr) if lens else p(text)
            errors.append({"text": text, "sample_counts": lens, "raises": None})
        except ValueError as e:
            errors.append({"text": text, "sample_counts": lens, "raises": "ValueError", "msg": str(e)})
    # collator (right and left padding)
    coll = []
    for side in ("right", "left"):
        p.tokenizer.padding_side = side
        samples = [p("Test with <|audio|>", audio=wave(0, sr), sampling_rate=sr),
                   p("Other longer text with <|audio|> more", audio=wave(1, 35 * sr), sampling_rate=sr)]
        for s in samples:
            s["input_ids"].squeeze_(0)
            s["attention_mask"].squeeze_(0)
        c = ultravox_processing.DataCollatorForSeq2SeqWithAudio(p.tokenizer)
        r = c([dict(s) for s in samples])
        d = tolist(r)
        d["audio_values_shape"] = list(r["audio_values"].shape)
        d["audio_values_sum"] = float(r["audio_values"].double().sum())
        coll.append({"padding_side": side, "out": d})
    p.tokenizer.padding_side = "right"
    json.dump({"cases": cases, "errors": errors, "collator": coll}, open(os.path.join(OUT, "processor_cases.json"), "w"),
              indent=1)
    print("processor cases:", len(cases), "errors:", len(errors))


def logmel_goldens():
    out = {}
    for n_mels, tag in ((80, "m80"), (128, "m128")):
        fe = transformers.WhisperFeatureExtractor(feature_size=n_mels)
        sr = 16000
        t = np.arange(2 * sr) / sr
        tone = (0.1 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
        tone[int(0.75 * len(tone)):] = 0
        waves = [wave(0, sr), tone, wave(2, sr // 2 + 33)]
        r = fe(waves, sampling_rate=sr, padding="longest", pad_to_multiple_of=160, truncation=False,
               return_attention_mask=True)
        out[f"{tag}_mel"] = np.asarray(r["input_features"], dtype=np.float32)
        out[f"{tag}_mask"] = np.asarray(r["attention_mask"], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "logmel_hf.npz"), **out)
    print("logmel goldens:", {k: v.shape for k, v in out.items()})


def projector_goldens():
    torch.manual_seed(0)
    out = {}
    for tag, ln_mid in (("v05", True), ("v04", False)):
        cfg = ultravox_config.UltravoxConfig(
            audio_config=dict(model_type="whisper", d_model=64, encoder_layers=1, encoder_attention_heads=2,
                              encoder_ffn_dim=128, num_mel_bins=80),
            text_config=dict(model_type="llama", hidden_size=96, num_hidden_layers=1, num_attention_heads=2,
                             intermediate_size=128, vocab_size=512),
            hidden_size=128, stack_factor=8, projector_ln_mid=ln_mid, norm_init=0.4)
        transformers.activations.ACT2FN["swiglu"] = ultravox_model.SwiGLU
        proj = ultravox_model.UltravoxProjector(cfg).float()
        with torch.no_grad():
            for n, p in proj.named_parameters():
                if "ln_" not in n:
                    p.copy_(torch.randn_like(p) * 0.05)
        x = torch.randn(3, 50, 64)
        with torch.no_grad():
            y = proj(x)
            st = ultravox_model.StackAudioFrames(8)(x)
        out[f"{tag}_in"] = x.numpy()
        out[f"{tag}_out"] = y.numpy()
        out[f"{tag}_stacked"] = st.numpy()
        for n, p in proj.state_dict().items():
            out[f"{tag}_w_{n}"] = p.numpy()
    # splice loop (UltravoxModel._prepare_audio_embeds / _audio_iter) executed verbatim on plain tensors
    emb = torch.randn(3, 40, 16)
    audio = torch.randn(4, 12, 16)
    start = torch.tensor([5, 20, 0, 28])
    tlen = torch.tensor([7, 12, 3, 12], dtype=torch.int32)
    abs_ = torch.tensor([2, 0, 2])
    got = emb.clone()
    for i_b, i_a in ultravox_model.UltravoxModel._audio_iter(None, abs_):
        s, n = start[i_a], tlen[i_a]
        got[i_b][s: s + n] = audio[i_a][:n]
    out.update(splice_emb=emb.numpy(), splice_audio=audio.numpy(), splice_start=start.numpy(), splice_len=tlen.numpy(),
               splice_abs=abs_.numpy(), splice_out=got.numpy())
    np.savez_compressed(os.path.join(OUT, "projector_ref.npz"), **out)
    print("projector goldens written")


def mask_goldens():
    """ModifiedWhisperEncoder.init_latency_mask (ref ultravox_model.py:834-863) run verbatim on a bare object."""
    enc = types.SimpleNamespace(max_context_length=3000)
    enc.register_buffer = lambda name, t, persistent=False: setattr(enc, name, t)
    ultravox_model.ModifiedWhisperEncoder.init_latency_mask(enc, 100, torch.float32)
    m = enc.audio_streaming_mask
    np.savez_compressed(os.path.join(OUT, "latency_mask.npz"), block=np.int64(100),
                        allowed=(m[0, 0, ::50, ::50] == 0).numpy())
    print("mask golden", tuple(m.shape))


CHAT_MESSAGES = [
    ("audio_qa", [{"role": "user", "content": "Listen to <|audio|> and answer briefly"},
                  {"role": "assistant", "content": "it says hello world again and again"}], 16000, "hello world"),
    ("system_audio", [{"role": "system", "content": "you are helpful"}, {"role": "user", "content": "Transcribe <|audio|>"},
                      {"role": "assistant", "content": "one two three four five six seven eight nine ten"}], 5 * 16000 + 123,
     "one two three"),
    ("text_only", [{"role": "user", "content": "What is two plus two"}, {"role": "assistant", "content": "four of course"}], 0, None),
    ("long_clip", [{"role": "user", "content": "<|audio|> summarise"}, {"role": "assistant", "content": "a b c d e f"}],
     35 * 16000, "long transcript words here"),
]


def dataproc_goldens():
    for name in ("librosa", "soundfile"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import importlib.util
    spec = importlib.util.spec_from_file_location("ultravox_data_sample_real", os.path.join(REF, "ultravox/data/data_sample.py"))
    ds = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ds)

    class _Base:                                   # stand-in for ultravox.data.Dataproc (ref datasets.py:592-616)
        def __init__(self, dataset):
            self._dataset = dataset

    fake = types.ModuleType("ultravox.data")
    fake.Dataproc, fake.VoiceSample = _Base, ds.VoiceSample
    fake.SizedIterableDataset, fake.Augmentation = object, object
    sys.modules["ultravox.data"] = fake
    import ultravox
    ultravox.data = fake
    from ultravox.model import ultravox_data_proc

    def chat(messages, tokenize=False, chat_template=None):
        return " ".join(f"<|start|> {m['role']} <|sep|> {m['content']} <|eot_id|>" for m in messages)

    proc = make_processor(80)
    proc.tokenizer.apply_chat_template = chat
    out = []
    for cname, messages, n, transcript in CHAT_MESSAGES:
        for mask in ("last_assistant", "after_audio", "all"):
            if mask == "after_audio" and n == 0:
                continue
            for alt, cap, infer in ((False, None, False), (True, None, False), (True, 3, False), (False, None, True)):
                audio = wave(7, n) if n else None
                if audio is not None and cname == "system_audio":
                    audio = (audio * 1000).astype(np.int16)          # exercises VoiceSample's integer normalisation
                sample = ds.VoiceSample([dict(m) for m in messages], audio, audio_transcript=transcript)
                dp = ultravox_data_proc.UltravoxDataproc(None, proc, ultravox_config.LossMaskType(mask), inference_mode=infer,
                                                         include_alt_fields=alt, max_response_tokens=cap)
                try:
                    r = dp._process(sample)
                except ValueError as e:               # e.g. inference_mode + last_assistant drops the turn that holds the audio
                    out.append({"name": cname, "n_samples": n, "int16": cname == "system_audio" and n > 0, "loss_mask_type": mask,
                                "include_alt_fields": alt, "max_response_tokens": cap, "inference_mode": infer,
                                "transcript": transcript, "raises": "ValueError", "msg": str(e)})
                    continue
                d = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in r.items() if k != "audio_values"}
                if "audio_values" in r:
                    d["audio_values_shape"] = list(r["audio_values"].shape)
                out.append({"name": cname, "n_samples": n, "int16": cname == "system_audio" and n > 0, "loss_mask_type": mask,
                            "include_alt_fields": alt, "max_response_tokens": cap, "inference_mode": infer,
                            "transcript": transcript, "out": d})
    json.dump({"messages": {c[0]: c[1] for c in CHAT_MESSAGES}, "cases": out}, open(os.path.join(OUT, "dataproc_cases.json"), "w"),
              indent=0)
    print("dataproc cases:", len(out))


if __name__ == "__main__":
    if "--dataproc-only" in sys.argv:
        dataproc_goldens()
        sys.exit(0)
    processor_cases()
    logmel_goldens()
    projector_goldens()
    mask_goldens()
    dataproc_goldens()
