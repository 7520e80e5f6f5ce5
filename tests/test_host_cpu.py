"""Host-side logic of the product package, no GPU: processor / collator integer path (bit-exact against fixtures made by
the reference class), config surface, C-ABI symbol table, error behaviour."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from ultravox_b200 import _lib
from ultravox_b200.config import LossConfig, LossFunction, UltravoxConfig, preset
from ultravox_b200.processing import DataCollatorForSeq2SeqWithAudio, MelSpec, UltravoxProcessor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


class StubTokenizer:
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
        import transformers
        L = max(len(f["input_ids"]) for f in features)
        out = {"input_ids": [], "attention_mask": []}
        for f in features:
            ids = list(map(int, f["input_ids"]))
            n = L - len(ids)
            left = self.padding_side == "left"
            out["input_ids"].append([self.pad_token_id] * n + ids if left else ids + [self.pad_token_id] * n)
            out["attention_mask"].append([0] * n + [1] * len(ids) if left else [1] * len(ids) + [0] * n)
        out.update({k: [f[k] for f in features] for k in features[0] if k not in ("input_ids", "attention_mask")})
        return transformers.BatchFeature(out, tensor_type=return_tensors)


def wave(i, n):
    return np.random.default_rng(1000 + i).standard_normal(n).astype(np.float32)


@pytest.fixture(scope="module")
def cases():
    return json.load(open(os.path.join(G, "processor_cases.json")))


@pytest.fixture()
def proc():
    return UltravoxProcessor(MelSpec(feature_size=80), StubTokenizer(), defer_mel=True)


def test_processor_integer_path_bit_exact(proc, cases):
    for c in cases["cases"]:
        audios = [wave(i, n) for i, n in enumerate(c["sample_counts"])]
        kw = dict(audios=audios, sampling_rate=16000, include_audio_num_chunks=True) if audios else {}
        got = proc(c["text"], **kw)
        exp = c["out"]
        for key in ("input_ids", "attention_mask", "audio_lens", "audio_token_len", "audio_token_start_idx",
                    "audio_batch_size", "audio_num_chunks"):
            if key in exp:
                assert got[key].tolist() == exp[key], (c["name"], key)
                if key == "audio_token_len":
                    assert got[key].dtype == torch.int32
                if key in ("audio_lens", "audio_token_start_idx", "input_ids"):
                    assert got[key].dtype == torch.int64
            else:
                assert key not in got
        if audios:
            L = got["audio_waveforms"].shape[1]
            assert L % 160 == 0 and L >= 320 and got["audio_waveforms"].dtype == torch.float32


def test_processor_errors(proc, cases):
    for e in cases["errors"]:
        audios = [wave(i, n) for i, n in enumerate(e["sample_counts"])]
        if e["raises"]:
            with pytest.raises(ValueError) as ei:
                proc(e["text"], audios=audios, sampling_rate=16000) if audios else proc(e["text"])
            assert str(ei.value) == e["msg"]
    with pytest.raises(ValueError):
        proc("x <|audio|>", audio=wave(0, 100), audios=[wave(0, 100)])
    with pytest.raises(ValueError):
        proc(["a", "b"])
    tok = StubTokenizer()
    tok.eos_token = None
    with pytest.raises(AssertionError):
        UltravoxProcessor(MelSpec(), tok)


def test_chunk_and_pad_audio_matches_reference_semantics(proc):
    """ref :153-215: continuation chunks are zero-padded to the context, first chunks keep the batch width."""
    mel = torch.arange(2 * 3 * 3500, dtype=torch.float32).view(2, 3, 3500)
    d = proc._chunk_and_pad_audio(mel, torch.tensor([3500, 1200]), include_audio_num_chunks=True)
    assert d["audio_values"].shape == (3, 3, 3000)
    assert d["audio_lens"].tolist() == [3000, 500, 1200] and d["audio_is_continuation"].tolist() == [False, True, False]
    assert torch.equal(d["audio_values"][1, :, :500], mel[0, :, 3000:]) and float(d["audio_values"][1, :, 500:].abs().sum()) == 0
    assert torch.equal(d["audio_values"][2], mel[1, :, :3000])
    assert d["audio_batch_size"].tolist() == [3] and d["audio_num_chunks"].tolist() == [2, 1]


def test_collator_left_padding_displacement(cases):
    sr = 16000
    for c in cases["collator"]:
        tok = StubTokenizer()
        tok.padding_side = c["padding_side"]
        p = UltravoxProcessor(MelSpec(feature_size=80), tok, defer_mel=True)
        samples = []
        for text, n, i in (("Test with <|audio|>", sr, 0), ("Other longer text with <|audio|> more", 35 * sr, 1)):
            s = dict(p(text, audio=wave(i, n), sampling_rate=sr))
            s.pop("audio_waveforms"), s.pop("audio_num_frames")
            # stand-in mel with the right frame width per chunk (the collator only pads / stacks it)
            widths = [3000 if len(s["audio_lens"]) > 1 else int(s["audio_lens"][0])] * len(s["audio_lens"])
            s["audio_values"] = torch.ones(len(widths), 80, widths[0])
            s["input_ids"], s["attention_mask"] = s["input_ids"][0], s["attention_mask"][0]
            samples.append(s)
        got = DataCollatorForSeq2SeqWithAudio(tok)(samples)
        exp = c["out"]
        for key in ("input_ids", "attention_mask", "audio_lens", "audio_token_len", "audio_token_start_idx",
                    "audio_batch_size"):
            assert got[key].tolist() == exp[key], (c["padding_side"], key)
        assert list(got["audio_values"].shape) == exp["audio_values_shape"]


def test_collator_deferred_mel_layout(cases):
    """the deferred-mel samples (the DataLoader-safe mode) go through the collator as they are: same index vectors as
    the reference fixtures, waveforms flattened per clip and zero-padded to the batch-longest multiple of the hop, each clip's
    own padded width recorded (``audio_pad_frames``), nothing left behind for ``tokenizer.pad`` to choke on."""
    sr = 16000
    for c in cases["collator"]:
        tok = StubTokenizer()
        tok.padding_side = c["padding_side"]
        p = UltravoxProcessor(MelSpec(feature_size=80), tok, defer_mel=True)
        samples = []
        for text, n, i in (("Test with <|audio|>", sr, 0), ("Other longer text with <|audio|> more", 35 * sr, 1)):
            s = dict(p(text, audio=wave(i, n), sampling_rate=sr))
            s["input_ids"], s["attention_mask"] = s["input_ids"][0], s["attention_mask"][0]
            samples.append(s)
        got = DataCollatorForSeq2SeqWithAudio(tok)(samples)
        exp = c["out"]
        for key in ("input_ids", "attention_mask", "audio_lens", "audio_token_len", "audio_token_start_idx", "audio_batch_size"):
            assert got[key].tolist() == exp[key], (c["padding_side"], key)
        assert "audio_values" not in got
        assert tuple(got["audio_waveforms"].shape) == (2, 35 * sr) and got["audio_waveforms"].dtype == torch.float32
        assert got["audio_num_frames"].tolist() == [100, 3500] and got["audio_pad_frames"].tolist() == [100, 3500]
        assert float(got["audio_waveforms"][0, sr:].abs().sum()) == 0.0
        assert np.array_equal(got["audio_waveforms"][0, :sr].numpy(), wave(0, sr))


def test_config_surface_round_trip(tmp_path):
    cfg = preset("micro", audio_latency_block_size=100)
    assert cfg.model_type == "ultravox" and cfg.stack_factor == 8 and cfg.projector_ln_mid is True
    assert cfg.vocab_size == cfg.text_config.vocab_size and cfg.initializer_range == cfg.text_config.initializer_range
    cfg.save_pretrained(tmp_path)
    back = UltravoxConfig.from_pretrained(tmp_path)
    assert back.to_dict()["audio_latency_block_size"] == 100
    assert back.audio_config.d_model == cfg.audio_config.d_model and back.text_config.hidden_size == cfg.text_config.hidden_size
    assert back.text_model_lora_config["r"] == 0
    d = UltravoxConfig().to_diff_dict()
    assert "_attn_implementation_autoset" not in d.get("text_config", {})
    assert LossConfig().loss_function == LossFunction.CrossEntropy and LossConfig(LossFunction.KL_Divergence).requires_alt_fields


def test_library_exports_every_declared_symbol():
    """include/uvx.h is the contract: every function it declares must be exported and bound."""
    header = open(os.path.join(ROOT, "include", "uvx.h")).read()
    declared = set(re.findall(r"\b(uvx_[a-z0-9_]+)\s*\(", header))
    declared -= {"uvx_gemm_args", "uvx_attn_args"}
    assert declared, "no declarations parsed"
    lib = _lib.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/uvx.h but not exported by libuvx.so"
        assert name in _lib.SIGNATURES, f"{name} not bound in ultravox_b200/_lib.py"
    assert lib.uvx_abi_version() == 1
    assert ctypes.sizeof(_lib.GemmArgs) == 8 * 18 + 4 * 3 + 4 + 16 + 24 + 8 + 5 * 8 + 8  # (+ flags / reserved) ... + w_tiled/rope_cols + 3 ptr + 2 int64 (round 2)


def test_host_only_mel_filter_table_is_bit_exact_with_hf():
    transformers = pytest.importorskip("transformers")
    lib = _lib.lib()
    for n in (80, 128):
        dense = np.zeros((201, n), np.float32)
        assert lib.uvx_debug_mel_filters(n, dense.ctypes.data) == 0
        hf = transformers.WhisperFeatureExtractor(feature_size=n).mel_filters.astype(np.float32)
        assert np.array_equal(dense, hf)


def test_no_cpu_fallback():
    from ultravox_b200 import ops
    with pytest.raises(_lib.UvxError):
        ops.logmel(torch.zeros(1, 320), 80)
    with pytest.raises(_lib.UvxError):
        ops.linear(torch.zeros(4, 64, dtype=torch.bfloat16), torch.zeros(64, 64, dtype=torch.bfloat16))
    assert lib_error_is_reported()


def lib_error_is_reported():
    lib = _lib.lib()
    rc = lib.uvx_debug_mel_filters(77, None)
    return rc == -1 and b"uvx_debug_mel_filters" in lib.uvx_last_error()


def test_dataproc_matches_reference_fixtures():
    """SURVEY 8f rank 4 (host feed): ``UltravoxDataproc._process`` + ``VoiceSample`` replay the fixtures written by the
    reference's own class (scripts/make_golden.py:dataproc_goldens; ref ultravox_data_proc.py:45-154, data_sample.py:88-100):
    loss-mask length for every LossMaskType, alt (text-only) twin, response truncation, inference mode, int16 audio."""
    from ultravox_b200.config import LossMaskType
    from ultravox_b200.data_proc import UltravoxDataproc, VoiceSample
    gold = json.load(open(os.path.join(G, "dataproc_cases.json")))

    def chat(messages, tokenize=False, chat_template=None):
        return " ".join(f"<|start|> {m['role']} <|sep|> {m['content']} <|eot_id|>" for m in messages)

    tok = StubTokenizer()
    tok.pad_token_id = tok.eos_token_id
    tok.apply_chat_template = chat
    proc = UltravoxProcessor(MelSpec(feature_size=80), tok, defer_mel=True)
    checked = raised = 0
    for c in gold["cases"]:
        audio = wave(7, c["n_samples"]) if c["n_samples"] else None
        if audio is not None and c["int16"]:
            audio = (audio * 1000).astype(np.int16)
        sample = VoiceSample([dict(m) for m in gold["messages"][c["name"]]], audio, audio_transcript=c["transcript"])
        if audio is not None:
            assert sample.audio.dtype == np.float32 and sample.audio.ndim == 1
        dp = UltravoxDataproc(None, proc, LossMaskType(c["loss_mask_type"]), inference_mode=c["inference_mode"],
                              include_alt_fields=c["include_alt_fields"], max_response_tokens=c["max_response_tokens"])
        if "raises" in c:
            with pytest.raises(ValueError) as ei:
                dp._process(sample)
            assert str(ei.value) == c["msg"]
            raised += 1
            continue
        r = dp._process(sample)
        for k, want in c["out"].items():
            if k == "audio_values_shape":
                continue
            got = r[k].tolist() if hasattr(r[k], "tolist") else r[k]
            assert got == want, (c["name"], c["loss_mask_type"], k)
        checked += 1
    assert checked >= 40 and raised >= 1
    # Dataproc is an iterable view of the wrapped dataset
    ds = [VoiceSample.from_prompt("What is two plus two"), VoiceSample.from_prompt_and_raw("Hear <|audio|>", wave(1, 16000), 16000)]
    ds[0].messages.append({"role": "assistant", "content": "four"})
    ds[1].messages.append({"role": "assistant", "content": "ok"})
    outs = list(UltravoxDataproc(ds, proc, LossMaskType.LAST_ASSISTANT))
    assert len(outs) == 2 and outs[1]["audio_token_len"].tolist() == [7] and outs[0]["labels"][-1] != -100
    with pytest.raises(AssertionError):
        VoiceSample([], np.zeros((2, 4), np.float32))
    with pytest.raises(AssertionError):
        VoiceSample([], np.zeros(4, np.uint8))


def test_pad_bounds_from_attention_mask():
    """Host half of the padding support: one contiguous run of ones per row -> (kv_start, kv_len); holes are rejected."""
    import torch
    from ultravox_b200.model import UltravoxModel
    f = UltravoxModel._pad_bounds
    assert f(None) == (None, None) and f(torch.ones(2, 5, dtype=torch.long)) == (None, None)
    right = torch.tensor([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1]])
    ks, kl = f(right)
    assert ks is None and kl.tolist() == [3, 5] and kl.dtype == torch.int32
    left = torch.tensor([[0, 0, 1, 1, 1], [1, 1, 1, 1, 1]])
    ks, kl = f(left)
    assert kl is None and ks.tolist() == [2, 0]
    both = torch.tensor([[0, 1, 1, 0, 0], [0, 0, 0, 0, 0]])
    ks, kl = f(both)
    assert ks.tolist() == [1, 0] and kl.tolist() == [3, 0]
    with pytest.raises(NotImplementedError):
        f(torch.tensor([[1, 0, 1, 1, 1]]))


def test_lora_checkpoint_names_and_merge():
    """SURVEY 8b state-dict row / 8f rank 3: PEFT-wrapped checkpoint names (base_model.model. / .base_layer. infixes,
    lora_A / lora_B) are folded into plain weights, W' = W + (alpha / r) B A, and the renaming round-trips."""
    import torch
    from ultravox_b200 import lora
    g = torch.Generator().manual_seed(0)
    d, r, alpha = 16, 4, 8.0
    plain = {"audio_tower.layers.0.self_attn.q_proj.weight": torch.randn(d, d, generator=g),
             "audio_tower.layers.0.self_attn.q_proj.bias": torch.randn(d, generator=g),
             "audio_tower.layers.0.self_attn.k_proj.weight": torch.randn(d, d, generator=g),
             "audio_tower.layers.0.self_attn.v_proj.weight": torch.randn(d, d, generator=g),
             "audio_tower.layer_norm.weight": torch.ones(d),
             "multi_modal_projector.linear_1.weight": torch.randn(8, d, generator=g)}
    wrapped = lora.to_lora_names(plain, "audio_tower", ["k_proj", "q_proj", "linear_k", "linear_q"])
    assert "audio_tower.base_model.model.layers.0.self_attn.q_proj.base_layer.weight" in wrapped
    assert "audio_tower.base_model.model.layers.0.self_attn.q_proj.base_layer.bias" in wrapped
    assert "audio_tower.base_model.model.layers.0.self_attn.v_proj.weight" in wrapped           # not a target module
    assert "audio_tower.base_model.model.layer_norm.weight" in wrapped and "multi_modal_projector.linear_1.weight" in wrapped
    assert lora.to_lora_names(wrapped, "audio_tower", ["q_proj"]) == wrapped                     # already wrapped: untouched
    assert not lora.has_lora_keys(plain) and lora.has_lora_keys(wrapped)
    A = {m: torch.randn(r, d, generator=g) for m in ("q_proj", "k_proj")}
    B = {m: torch.randn(d, r, generator=g) for m in ("q_proj", "k_proj")}
    for m in A:
        stem = f"audio_tower.base_model.model.layers.0.self_attn.{m}"
        wrapped[f"{stem}.lora_A.default.weight"], wrapped[f"{stem}.lora_B.default.weight"] = A[m], B[m]
    scal = {"audio_tower": lora.lora_scaling({"r": r, "lora_alpha": alpha}), "language_model": lora.lora_scaling({"r": 0})}
    assert scal == {"audio_tower": 2.0, "language_model": 0.0}
    merged = lora.merge_lora_state_dict(wrapped, scal)
    assert set(merged) == set(plain)
    x = torch.randn(5, d, generator=g)
    for m in A:
        k = f"audio_tower.layers.0.self_attn.{m}.weight"
        want = x @ plain[k].T + (alpha / r) * (x @ A[m].T) @ B[m].T      # peft LoRA forward: base(x) + scaling * B(A(x))
        assert torch.allclose(x @ merged[k].T, want, atol=1e-4)
    assert torch.equal(merged["audio_tower.layers.0.self_attn.v_proj.weight"], plain["audio_tower.layers.0.self_attn.v_proj.weight"])
    assert torch.equal(merged["audio_tower.layers.0.self_attn.q_proj.bias"], plain["audio_tower.layers.0.self_attn.q_proj.bias"])
    with pytest.raises(ValueError):
        lora.merge_lora_state_dict(wrapped, {"audio_tower": 0.0})
    broken = dict(wrapped)
    del broken["audio_tower.base_model.model.layers.0.self_attn.q_proj.lora_B.default.weight"]
    with pytest.raises(KeyError):
        lora.merge_lora_state_dict(broken, scal)


def test_model_loads_lora_wrapped_checkpoint():
    """``UltravoxModel.load_state_dict`` takes a PEFT-named checkpoint (LoRA r=4 on the Whisper q/k projections) and
    ends up with merged weights in the fused q|k|v storage the GEMM reads (no GPU needed: parameter containers only)."""
    import torch
    from ultravox_b200 import lora
    from ultravox_b200.config import preset
    from ultravox_b200.model import UltravoxModel
    cfg = preset("micro")
    cfg.audio_model_lora_config = {"r": 4, "lora_alpha": 8, "target_modules": ["k_proj", "q_proj"]}
    m = UltravoxModel(cfg, device="cpu")
    g = torch.Generator().manual_seed(0)
    sd = {k: (torch.randn(v.shape, generator=g) * 0.02).to(v.dtype) for k, v in m.state_dict().items()}
    wrapped = lora.to_lora_names(sd, "audio_tower", ["k_proj", "q_proj"])
    d = sd["audio_tower.layers.0.self_attn.q_proj.weight"].shape[0]
    ab = {}
    for layer in range(cfg.audio_config.encoder_layers):
        for mod in ("q_proj", "k_proj"):
            stem = f"audio_tower.base_model.model.layers.{layer}.self_attn.{mod}"
            ab[stem] = (torch.randn(4, d, generator=g).to(torch.bfloat16), torch.randn(d, 4, generator=g).to(torch.bfloat16))
            wrapped[stem + ".lora_A.default.weight"], wrapped[stem + ".lora_B.default.weight"] = ab[stem]
    res = m.load_state_dict(wrapped)
    assert not res.missing_keys and not res.unexpected_keys
    now = m.state_dict()
    for stem, (a, b) in ab.items():
        k = lora.plain_name(stem + ".base_layer.weight")
        want = sd[k].float() + 2.0 * (b.float() @ a.float())
        assert ((now[k].float() - want).norm() / want.norm()).item() < 4e-3          # one bf16 rounding of the merged weight
    att = m.audio_tower.layers[0].self_attn
    assert torch.equal(att.qkv_w[:d], now["audio_tower.layers.0.self_attn.q_proj.weight"])   # views of the fused GEMM operand
    assert torch.equal(now["audio_tower.layers.0.self_attn.v_proj.weight"], sd["audio_tower.layers.0.self_attn.v_proj.weight"])
    cfg0 = preset("micro")
    with pytest.raises(ValueError):                                                    # LoRA weights but r = 0 in the config
        UltravoxModel(cfg0, device="cpu").load_state_dict(wrapped)


def test_local_inference_host_logic():
    """Host half of the LocalInference mirror (ref inference/infer.py:52-124, base.py): conversation bookkeeping, thinking
    post-processing, resampling length, the token queue; generation itself is covered by the GPU tests."""
    import torch
    from ultravox_b200.data_proc import VoiceSample
    from ultravox_b200.inference import (InferenceChunk, InferenceStats, LocalInference, VoiceInference, VoiceOutput, _TokenQueue,
                                         resample_to_16k)

    class FakeModel:
        device = torch.device("cpu")

        def eval(self):
            return self

    tok = StubTokenizer()
    tok.padding_side = "left"
    inf = LocalInference(FakeModel(), object(), tok, conversation_mode=True)
    with pytest.raises(ValueError):
        inf._get_sample_with_past(None)                               # nothing to continue from
    msgs = [{"role": "user", "content": "hear <|audio|> now"}]
    past = inf._build_past_messages(msgs, 3, "fine")
    assert past == [{"role": "user", "content": "hear " + tok.eos_token * 3 + " now"}, {"role": "assistant", "content": "fine"}]
    assert msgs[0]["content"] == "hear <|audio|> now"                  # the query itself is not edited in place
    with pytest.raises(ValueError):
        inf._build_past_messages([{"role": "user", "content": "<|audio|> <|audio|>"}], 3, "x")
    assert inf._build_past_messages([{"role": "user", "content": "text only"}], 0, "ok")[-1]["content"] == "ok"
    inf.update_conversation(past, "cache")
    s = inf._get_sample_with_past(VoiceSample.from_prompt("next"))
    assert [m["role"] for m in s.messages] == ["user", "assistant", "user"] and inf.past_key_values == "cache"
    assert inf._get_sample_with_past(None).messages == past
    inf.update_conversation()
    assert inf.past_messages == [] and inf.past_key_values is None
    # thinking post-processing
    assert inf._postprocess_response("plain") == ("plain", None)
    th = LocalInference(FakeModel(), object(), tok, enable_thinking=True, thinking_regex=r"<think>(.*?)</think>")
    assert th._postprocess_response("<think> a b </think> answer") == ("answer", "a b")
    with pytest.raises(ValueError):
        th._postprocess_response("no thoughts here")
    with pytest.raises(ValueError):
        LocalInference(FakeModel(), object(), tok, enable_thinking=True)._postprocess_response("x")
    with pytest.raises(ValueError):
        LocalInference(FakeModel(), object(), tok, dtype=torch.float16)
    tok_r = StubTokenizer()                                            # padding_side "right" is refused like the reference
    with pytest.raises(AssertionError):
        LocalInference(FakeModel(), object(), tok_r)
    # resampling: 48 kHz -> 16 kHz keeps duration (ref infer_test.py:112-132 pins the resulting frame count)
    x = np.sin(2 * np.pi * 440 * np.arange(48000) / 48000).astype(np.float32)
    y = resample_to_16k(x, 48000)
    assert y.dtype == np.float32 and len(y) == 16000
    ref = np.sin(2 * np.pi * 440 * np.arange(16000) / 16000)
    assert np.abs(y[200:-200] - ref[200:-200]).max() < 2e-3
    assert resample_to_16k(x[:16000], 16000) is not None and len(resample_to_16k(x[:44100], 44100)) == 16000
    # streamer protocol: the prompt is skipped, tokens flow, end() terminates the iterator
    q = _TokenQueue()
    q.put(torch.tensor([[1, 2, 3]]))
    q.put(torch.tensor([7]))
    q.put(torch.tensor([9]))
    q.end()
    assert list(q) == [7, 9]

    class Once(VoiceInference):                                        # base-class fallbacks
        def infer(self, sample, max_tokens=None, temperature=None):
            return VoiceOutput("hi", 3, 1)

    o = Once()
    assert [v.text for v in o.infer_batch([None, None])] == ["hi", "hi"]
    out = list(o.infer_stream(None))
    assert out == [InferenceChunk("hi"), InferenceStats(3, 1)]


def test_transformers_registration_and_repetition_penalty():
    """SURVEY 8b registration row (ref ultravox_model.py:997-1003, ultravox_processing.py:385-387, ultravox_pipeline.py:128-133)
    and the pipeline's default logits processor (hf RepetitionPenaltyLogitsProcessor) - bit-exact on CPU."""
    import torch
    import transformers
    from transformers.activations import ACT2FN
    from transformers.generation.logits_process import RepetitionPenaltyLogitsProcessor
    from transformers.models.auto.modeling_auto import MODEL_MAPPING
    from transformers.models.auto.processing_auto import PROCESSOR_MAPPING
    from ultravox_b200 import model, pipeline
    from ultravox_b200.config import UltravoxConfig
    assert isinstance(transformers.AutoConfig.for_model("ultravox"), UltravoxConfig)
    assert MODEL_MAPPING[UltravoxConfig] is model.UltravoxModel
    assert PROCESSOR_MAPPING[UltravoxConfig] is UltravoxProcessor
    assert isinstance(ACT2FN["swiglu"], model.SwiGLU)
    task = transformers.pipelines.PIPELINE_REGISTRY.check_task("ultravox-pipeline")
    assert task[1]["impl"] is pipeline.UltravoxPipeline
    g = torch.Generator().manual_seed(0)
    for pen in (1.0, 1.1, 2.5):
        lg, seq = torch.randn(3, 50, generator=g), torch.randint(0, 50, (3, 9), generator=g)
        assert torch.equal(model.apply_repetition_penalty(lg.clone(), seq, pen), RepetitionPenaltyLogitsProcessor(pen)(seq, lg.clone())
                           if pen != 1.0 else lg)


def test_pipeline_host_stages():
    """ref ultravox_pipeline.py:52-126: parameter split, audio dtype normalisation, prompt / turns handling, terminators, slicing."""
    import torch
    from ultravox_b200.pipeline import UltravoxPipeline

    def chat(turns, add_generation_prompt=True, tokenize=False, **kw):
        return " ".join(f"<s> {m['role']} : {m['content']}" for m in turns) + (" <s> assistant :" if add_generation_prompt else "")

    tok = StubTokenizer()
    tok.pad_token_id = tok.eos_token_id
    tok.apply_chat_template = chat
    tok.added_tokens_encoder = {"<|eot_id|>": 128009}
    tok.convert_tokens_to_ids = lambda t: tok.added_tokens_encoder[t]
    tok.decode = lambda ids, skip_special_tokens=True: " ".join(str(int(i)) for i in ids)
    calls = {}

    class FakeModel:
        device = torch.device("cpu")

        def generate(self, **kw):
            calls.update(kw)
            return torch.cat([kw["input_ids"], torch.tensor([[11, 12, 13]])], dim=1)

    proc = UltravoxProcessor(MelSpec(feature_size=80), tok, defer_mel=True)
    pipe = UltravoxPipeline(FakeModel(), tokenizer=tok, processor=proc)
    assert pipe._sanitize_parameters(temperature=0.5, max_new_tokens=7, foo=1) == ({}, {"temperature": 0.5, "max_new_tokens": 7}, {})
    a16 = (wave(0, 16000) * 2000).astype(np.int16)
    out = pipe.preprocess({"audio": a16, "sampling_rate": 16000})
    assert out["audio_token_len"].tolist() == [7] and out["audio_waveforms"].dtype == torch.float32
    assert float(out["audio_waveforms"].abs().max()) <= 1.0                                  # int16 / 32768
    turns = [{"role": "system", "content": "be brief"}]
    out2 = pipe.preprocess({"audio": wave(0, 16000).astype(np.float64), "turns": turns, "prompt": "what is said", "sampling_rate": 16000})
    assert turns[-1] == {"role": "user", "content": "what is said <|audio|>"} and out2["audio_token_len"].tolist() == [7]
    keep = [{"role": "user", "content": "about <|audio|> please"}]
    pipe.preprocess({"audio": wave(1, 32000), "turns": keep, "sampling_rate": 16000})
    assert len(keep) == 1                                                                    # the last turn is the user's: used as is
    text = pipe({"audio": wave(0, 16000), "sampling_rate": 16000}, max_new_tokens=3)
    assert text == "11 12 13" and calls["repetition_penalty"] == 1.1 and calls["do_sample"] is False
    assert calls["eos_token_id"] == [128009, 128009] and calls["max_new_tokens"] == 3
    with pytest.raises(ValueError):
        UltravoxPipeline(FakeModel())
