"""End-to-end parity of the CUDA path (through the C ABI) against the fp32 CPU oracle on shared seeded weights.

Stage tolerances follow SURVEY.md section 7: bf16 storage between kernels means the end-to-end error against an fp32
oracle is a few 1e-3 .. 1e-2 relative (HF's own bf16 forward shows 7e-3 after one layer, 1.3e-2 after 32); index / splice
paths are bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def wave(i, n):
    return np.random.default_rng(1000 + i).standard_normal(n).astype(np.float32)


def build(name="micro", **kw):
    from ultravox_b200.config import preset
    from ultravox_b200.model import UltravoxModel
    from oracle import model as om
    cfg = preset(name, **kw)
    model = UltravoxModel(cfg, device="cuda").init_random_(seed=42)
    return cfg, model, om.state_dict_fp32(model), om.shapes_from_config(cfg)


def make_batch(cfg, lens_samples, text_pre=8, text_post=5, seed=7):
    """One sequence per clip (B = len(lens)), layout of ref infer_test.py:97-109: 8 ids, placeholders, 5 ids."""
    from oracle import logmel as ol
    waves = [wave(i, n) for i, n in enumerate(lens_samples)]
    padded, frames = ol.pad_batch(waves)
    g = torch.Generator().manual_seed(seed)
    tok = [int(-(-int(f) // 16)) for f in frames]
    S = text_pre + max(tok) + text_post
    ids = torch.randint(0, cfg.vocab_size, (len(waves), S), generator=g)
    batch = dict(input_ids=ids, audio_token_start_idx=torch.tensor([text_pre] * len(waves)),
                 audio_lens=torch.tensor([int(f) for f in frames]), audio_token_len=torch.tensor(tok, dtype=torch.int32),
                 audio_batch_size=torch.ones(len(waves), dtype=torch.int64))
    return padded, batch


def test_state_dict_keys_match_reference_names():
    cfg, model, sd, sh = build()
    keys = set(sd)
    for k in ("audio_tower.conv1.weight", "audio_tower.conv2.bias", "audio_tower.embed_positions.weight",
              "audio_tower.layers.0.self_attn.q_proj.weight", "audio_tower.layers.0.self_attn.q_proj.bias",
              "audio_tower.layers.1.self_attn.k_proj.weight", "audio_tower.layers.0.self_attn.out_proj.bias",
              "audio_tower.layers.0.self_attn_layer_norm.weight", "audio_tower.layers.0.fc1.weight",
              "audio_tower.layers.0.final_layer_norm.bias", "audio_tower.layer_norm.weight",
              "multi_modal_projector.ln_pre.weight", "multi_modal_projector.linear_1.weight",
              "multi_modal_projector.ln_mid.weight", "multi_modal_projector.linear_2.weight",
              "language_model.model.embed_tokens.weight", "language_model.model.layers.0.self_attn.q_proj.weight",
              "language_model.model.layers.1.self_attn.o_proj.weight", "language_model.model.layers.0.mlp.gate_proj.weight",
              "language_model.model.layers.0.mlp.down_proj.weight", "language_model.model.layers.0.input_layernorm.weight",
              "language_model.model.layers.0.post_attention_layernorm.weight", "language_model.model.norm.weight",
              "language_model.lm_head.weight"):
        assert k in keys, k
    assert "audio_tower.layers.0.self_attn.k_proj.bias" not in keys
    assert not any("qkv" in k or "gate_up" in k for k in keys)
    # fused storage really is shared: writing through the named parameter changes the fused weight
    l0 = model.language_model.model.layers[0].self_attn
    l0.k_proj.weight.data.fill_(0.5)
    nq = cfg.text_config.num_attention_heads * model.language_model.head_dim
    assert float(l0.qkv_w[nq:nq + 4].float().mean()) == 0.5
    assert set(model.diff_state_dict().keys()) == {k for k in keys if k.startswith("multi_modal_projector.")}


@pytest.mark.parametrize("lens", [[16000], [16000 * 3, 16000 + 77, 5000]])
def test_encoder_projector_stages(lens):
    from oracle import logmel as ol, model as om
    from ultravox_b200 import ops
    cfg, model, sd, sh = build()
    padded, batch = make_batch(cfg, lens)
    mel_ref = torch.from_numpy(ol.log_mel(padded, sh.n_mels))
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    assert float((mel.cpu() - mel_ref).abs().max()) < 2e-3
    # feed the oracle the bf16-rounded mel the GPU path actually consumes -> isolates encoder/projector error
    mel_b = mel.cpu().to(torch.bfloat16).float()
    enc_ref = om.whisper_encoder(sd, sh, mel_b, batch["audio_lens"])
    tm = ops.mel_to_timemajor(mel)
    enc = model.encode_audio(tm, batch["audio_lens"])
    for i, n in enumerate(batch["audio_lens"].tolist()):
        valid = (n - 1) // 2 + 1
        assert rel(enc[i, :valid], enc_ref[i, :valid]) < 1.5e-2, i
    aud_ref = om.projector(sd, sh, enc_ref)
    aud = model.project_audio(enc)
    for i, n in enumerate(batch["audio_token_len"].tolist()):
        assert rel(aud[i, :n], aud_ref[i, :n]) < 2e-2, i


def test_forward_matches_oracle_ragged_batch():
    from oracle import logmel as ol, model as om
    from ultravox_b200 import ops
    cfg, model, sd, sh = build()
    padded, batch = make_batch(cfg, [16000 * 2, 16000 + 77, 9000])
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    labels = batch["input_ids"].clone()
    labels[:, :-5] = -100
    out = model(audio_values=mel, labels=labels.cuda(), **{k: v.cuda() for k, v in batch.items()})
    st = {}
    ref_logits, ref_loss = om.forward(sd, sh, batch["input_ids"], mel.cpu().to(torch.bfloat16).float(),
                                      batch["audio_token_start_idx"], batch["audio_lens"], batch["audio_token_len"],
                                      batch["audio_batch_size"], labels=labels, stages=st)
    assert out.logits.shape == ref_logits.shape and out.logits.dtype == torch.float32
    r = rel(out.logits, ref_logits)
    agree = float((out.logits.cpu().argmax(-1) == ref_logits.argmax(-1)).float().mean())
    assert r < 3e-2 and agree > 0.9, (r, agree)
    assert abs(float(out.loss) - float(ref_loss)) < 3e-2 * max(1.0, abs(float(ref_loss)))


def test_splice_rows_bit_exact_and_text_rows_are_table_rows():
    from oracle import logmel as ol
    from ultravox_b200 import ops
    cfg, model, sd, sh = build()
    padded, batch = make_batch(cfg, [16000, 16000 * 2])
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    emb = model._prepare_audio_embeds(batch["input_ids"].cuda(), mel, batch["audio_token_start_idx"], batch["audio_lens"],
                                      batch["audio_token_len"], batch["audio_batch_size"])
    aud = model.project_audio(model.encode_audio(ops.mel_to_timemajor(mel), batch["audio_lens"]))
    table = model.language_model.model.embed_tokens.weight
    for b in range(2):
        s, n = int(batch["audio_token_start_idx"][b]), int(batch["audio_token_len"][b])
        assert torch.equal(emb[b, s:s + n], aud[b, :n])
        ids = batch["input_ids"][b].cuda()
        assert torch.equal(emb[b, :s], table[ids[:s]]) and torch.equal(emb[b, s + n:], table[ids[s + n:]])


def test_text_only_and_right_padding_and_errors():
    from oracle import model as om
    cfg, model, sd, sh = build()
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, cfg.vocab_size, (2, 19), generator=g)
    am = torch.ones(2, 19, dtype=torch.long)
    am[1, 13:] = 0
    out = model(ids.cuda(), attention_mask=am.cuda())
    ref = om.llama_forward(sd, sh, sd["language_model.model.embed_tokens.weight"][ids], attention_mask=am)
    assert rel(out.logits[0], ref[0]) < 2e-2 and rel(out.logits[1, :13], ref[1, :13]) < 2e-2
    hole = am.clone()
    hole[0, 5] = 0                                     # interior hole: not a padding pattern
    with pytest.raises(NotImplementedError):
        model(ids.cuda(), attention_mask=hole.cuda())
    with pytest.raises(AssertionError):
        model(ids.cuda(), audio_values=torch.zeros(1, 80, 100).cuda())
    with pytest.raises(ValueError):
        model.encode_audio(torch.zeros(1, 3003, 80, dtype=torch.bfloat16, device="cuda"), None)


def test_left_padded_batch_forward_and_generate():
    """left-padded batches (the reference batches inference prompts with padding_side="left", ref
    infer.py:155-180, ultravox_processing.py:53-63).  forward: keys in the padding are masked (oracle = HF Llama with the
    same attention_mask, arange positions, hf:modeling_llama.py:394-397); generate: mask-derived position_ids
    (hf:generation/utils.py:707-729), so every row must continue exactly like its own unpadded prompt."""
    from oracle import model as om
    cfg, model, sd, sh = build()
    g = torch.Generator().manual_seed(3)
    S, pads = 21, [0, 6, 13]
    ids = torch.randint(0, cfg.vocab_size, (3, S), generator=g)
    am = torch.ones(3, S, dtype=torch.long)
    for b, pd in enumerate(pads):
        am[b, :pd] = 0
    table = sd["language_model.model.embed_tokens.weight"]
    out = model(ids.cuda(), attention_mask=am.cuda())
    ref = om.llama_forward(sd, sh, table[ids], attention_mask=am)
    for b, pd in enumerate(pads):
        assert rel(out.logits[b, pd:], ref[b, pd:]) < 2e-2
    # mask-derived positions: row b of the padded batch == the unpadded prompt on its own (oracle and CUDA path)
    pos = (am.cumsum(-1) - 1).clamp_min(0)
    out_p = model(ids.cuda(), attention_mask=am.cuda(), position_ids=pos.cuda())
    ref_p = om.llama_forward(sd, sh, table[ids], attention_mask=am, position_ids=pos)
    for b, pd in enumerate(pads):
        assert rel(out_p.logits[b, pd:], ref_p[b, pd:]) < 2e-2
        solo = om.llama_forward(sd, sh, table[ids[b:b + 1, pd:]])
        assert rel(out_p.logits[b, pd:], solo[0]) < 2e-2
    # generation: batch of left-padded prompts vs each prompt alone (same kernels) - logits of every step must agree;
    # tokens may only differ where the top-2 margin is inside bf16 noise, so compare through teacher forcing
    seq = model.generate(ids.cuda(), attention_mask=am.cuda(), max_new_tokens=5)
    assert seq.shape == (3, S + 5) and torch.equal(seq[:, :S].cpu(), ids)
    for b, pd in enumerate(pads):
        full = seq[b:b + 1, pd:]                                   # the row's own tokens, padding stripped
        lg = model(full).logits[0].float()                         # one unpadded re-forward of the generated sequence
        for t in range(5):
            row = lg[S - pd - 1 + t]
            top = row.topk(5).indices.tolist()
            assert int(seq[b, S + t]) in top, (b, t)
        single = model.generate(ids[b:b + 1, pd:].cuda(), max_new_tokens=5)
        agree = (single[0, S - pd:] == seq[b, S:]).float().mean().item()
        assert agree >= 0.6, (b, agree)
    with pytest.raises(NotImplementedError):
        model.generate(ids.cuda(), attention_mask=am.flip(1).cuda(), max_new_tokens=2)


def test_conversation_kv_reuse_two_turns():
    """SURVEY 8f rank 2 (conversation mode, ref infer.py:126-148): turn 2 passes the cache returned by turn 1 and only the
    new suffix is prefilled; the continuation must match a cache-less generation over the whole conversation."""
    from ultravox_b200 import ops
    cfg, model, sd, sh = build()
    padded, batch = make_batch(cfg, [16000])
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    kw = {k: v.cuda() for k, v in batch.items()}
    out1 = model.generate(audio_values=mel, max_new_tokens=4, return_dict_in_generate=True, **kw)
    S1 = batch["input_ids"].shape[1]
    cache = out1.past_key_values
    assert out1.sequences.shape == (1, S1 + 4) and cache.length == S1 + 3      # the last new token is not in the cache yet
    g = torch.Generator().manual_seed(9)
    turn2 = torch.randint(0, cfg.vocab_size, (1, 6), generator=g).cuda()
    ids2 = torch.cat([out1.sequences, turn2], dim=1)
    kw2 = dict(kw, input_ids=ids2)
    out2 = model.generate(audio_values=mel, max_new_tokens=5, past_key_values=cache, return_dict_in_generate=True, **kw2)
    S2 = ids2.shape[1]
    assert out2.sequences.shape == (1, S2 + 5) and torch.equal(out2.sequences[:, :S2], ids2)
    assert out2.past_key_values.length == S2 + 4 and out2.past_key_values.capacity >= S2 + 5
    # cache-less reference over the whole conversation (same kernels): teacher-forced logits must rank every token on top
    full = model(out2.sequences[:, :-1], audio_values=mel, **{k: v for k, v in kw.items() if k != "input_ids"}).logits[0].float()
    for t in range(5):
        assert int(out2.sequences[0, S2 + t]) in full[S2 - 1 + t].topk(5).indices.tolist(), t
    fresh = model.generate(audio_values=mel, max_new_tokens=5, **kw2)
    assert (fresh[0, S2:] == out2.sequences[0, S2:]).float().mean().item() >= 0.6
    with pytest.raises(ValueError):
        model.generate(audio_values=mel, max_new_tokens=2, past_key_values=out2.past_key_values, **kw)   # prompt shorter than cache


def test_generate_greedy_matches_stepwise_oracle():
    """KV-cache greedy decode against the fp32 oracle stepped token by token (the oracle re-forwards the whole sequence for every
    new token, no cache).  Random-init logits are nearly tied at the top, so the oracle is teacher-forced with the CUDA tokens:
    each CUDA token must be in the oracle's top-5 with a logit gap to its argmax inside the end-to-end tolerance."""
    from oracle import logmel as ol, model as om
    from ultravox_b200 import ops
    cfg, model, sd, sh = build()
    padded, batch = make_batch(cfg, [16000])
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    n_new = 6
    seq = model.generate(audio_values=mel, max_new_tokens=n_new, **{k: v.cuda() for k, v in batch.items()})
    S = batch["input_ids"].shape[1]
    assert seq.shape == (1, S + n_new) and torch.equal(seq[:, :S].cpu(), batch["input_ids"])
    st = {}
    om.forward(sd, sh, batch["input_ids"], mel.cpu().to(torch.bfloat16).float(), batch["audio_token_start_idx"], batch["audio_lens"],
               batch["audio_token_len"], batch["audio_batch_size"], last_only=True, stages=st)
    cur = st["inputs_embeds"]
    table = sd["language_model.model.embed_tokens.weight"]
    exact = 0
    for t in range(n_new):
        ref = om.llama_forward(sd, sh, cur, last_only=True).view(-1)
        tok = int(seq[0, S + t])
        assert tok in ref.topk(5).indices.tolist(), t
        assert float(ref.max() - ref[tok]) < 3e-2 * float(ref.abs().max()), t
        exact += int(tok == int(ref.argmax()))
        cur = torch.cat([cur, table[tok][None, None]], dim=1)
    assert exact >= n_new - 2, exact


def test_generate_cache_decode_matches_cacheless_reforward():
    """Self-consistency of the CUDA path: KV-cache decode == a full re-forward of the same kernels without a cache."""
    from ultravox_b200 import ops
    cfg, model, sd, sh = build()
    padded, batch = make_batch(cfg, [16000])
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    seq = model.generate(audio_values=mel, max_new_tokens=4, **{k: v.cuda() for k, v in batch.items()})
    S = batch["input_ids"].shape[1]
    emb = model._prepare_audio_embeds(batch["input_ids"].cuda(), mel, batch["audio_token_start_idx"], batch["audio_lens"],
                                      batch["audio_token_len"], batch["audio_batch_size"])
    table = model.language_model.model.embed_tokens.weight
    cur = emb
    for t in range(4):
        hidden = model.llama_hidden(cur.clone())
        tok = ops.argmax(ops.lm_head(hidden[:, -1, :], model.language_model.lm_head.weight))
        assert int(tok) == int(seq[0, S + t]), t
        cur = torch.cat([cur, table[tok][None]], dim=1)


def test_decode_engine_ignores_cache_tail_contents():
    """the static KV cache is torch.empty; rows past a stream's length must never reach the P.V product
    (a masked probability of 0 times NaN is NaN).  Same tokens with the cache pre-filled with NaN."""
    from ultravox_b200 import ops
    from ultravox_b200.engine import DecodeEngine
    cfg, model, sd, sh = build()
    padded, batch = make_batch(cfg, [16000])
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    emb = model._prepare_audio_embeds(batch["input_ids"].cuda(), mel, batch["audio_token_start_idx"], batch["audio_lens"],
                                      batch["audio_token_len"], batch["audio_batch_size"])
    runs = []
    for poison in (False, True):
        de = DecodeEngine(model, 1, 150)
        de.cache.k.fill_(float("nan") if poison else 0.0)
        de.cache.v.fill_(float("nan") if poison else 0.0)
        toks = [int(de.prefill(emb.clone())[0])]
        for _ in range(5):
            toks.append(int(de.step().view(-1)[0]))
        assert bool(torch.isfinite(de.logits).all())
        runs.append(toks)
    assert runs[0] == runs[1]


def test_generate_sampling_eos_and_deferred_waveforms():
    from ultravox_b200 import ops
    cfg, model, sd, sh = build()
    padded, batch = make_batch(cfg, [16000])
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    kw = {k: v.cuda() for k, v in batch.items()}
    S = batch["input_ids"].shape[1]
    greedy = model.generate(audio_values=mel, max_new_tokens=12, **kw)
    # (1) sampling: reproducible for a seeded generator, every token inside the top-k of the teacher-forced distribution
    g = torch.Generator(device="cuda").manual_seed(11)
    s1 = model.generate(audio_values=mel, max_new_tokens=12, do_sample=True, temperature=0.9, top_k=20, generator=g, **kw)
    g = torch.Generator(device="cuda").manual_seed(11)
    s2 = model.generate(audio_values=mel, max_new_tokens=12, do_sample=True, temperature=0.9, top_k=20, generator=g, **kw)
    assert torch.equal(s1, s2) and s1.shape == (1, S + 12)
    full = model(s1[:, :-1], audio_values=mel, **{k: v for k, v in kw.items() if k != "input_ids"}).logits[0].float()
    for t in range(12):
        assert int(s1[0, S + t]) in full[S - 1 + t].topk(24).indices.tolist(), t
    g3 = torch.Generator(device="cuda").manual_seed(12)
    s3 = model.generate(audio_values=mel, max_new_tokens=12, do_sample=True, temperature=0.9, top_k=20, generator=g3, **kw)
    assert not torch.equal(s1, s3) or not torch.equal(s1, greedy)
    # (2) EOS: generation stops right after the step at which the row finished (HF semantics), without a streamer
    eos = int(greedy[0, S + 2])
    first = (greedy[0, S:] == eos).nonzero()[0].item()
    out = model.generate(audio_values=mel, max_new_tokens=12, eos_token_id=[eos], **kw)
    assert out.shape == (1, S + first + 1) and torch.equal(out[0], greedy[0, :S + first + 1])
    # (3) deferred mel: raw waveform in, same tokens as the precomputed mel
    wv = torch.from_numpy(padded)
    out_w = model.generate(max_new_tokens=6, audio_waveforms=wv, audio_num_frames=torch.tensor([100]), **kw)
    assert torch.equal(out_w, greedy[:, :S + 6])
    with pytest.raises(TypeError):
        model.generate(audio_values=mel, max_new_tokens=2, audio_wavforms=wv, **kw)          # misspelt tensor kwarg is not swallowed
    # (4) repetition penalty through the graph == eager step-by-step application of the HF rule
    pen = model.generate(audio_values=mel, max_new_tokens=8, repetition_penalty=1.3, **kw)
    pen_eager = model.generate(audio_values=mel, max_new_tokens=8, repetition_penalty=1.3, use_graph=False, **kw)
    assert torch.equal(pen, pen_eager)


def test_prefill_engine_graph_matches_eager():
    from ultravox_b200.engine import PrefillEngine
    cfg, model, sd, sh = build()
    padded, batch = make_batch(cfg, [16000])
    eng = PrefillEngine(model, 16000, batch["input_ids"], batch["audio_token_start_idx"], batch["audio_token_len"],
                        batch["audio_batch_size"])
    assert eng.launches_per_step > 20
    host = torch.from_numpy(padded).pin_memory()
    tok = eng.run_e2e(host).clone()
    from ultravox_b200 import ops
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    out = model(audio_values=mel, logits_to_keep=1, **{k: v.cuda() for k, v in batch.items()})
    assert int(tok[0]) == int(out.logits.view(1, -1).argmax(-1))
    assert torch.allclose(eng.logits, out.logits.view(1, -1), rtol=1e-5, atol=1e-5)
    host2 = torch.from_numpy(wave(5, 16000)[None]).pin_memory()
    t2 = eng.run_e2e(host2).clone()
    eng.wave.copy_(host2)
    assert int(eng.run()[0]) == int(t2[0])


class _Tok:
    eos_token = "<|eot_id|>"
    eos_token_id = 1000
    pad_token_id = None
    padding_side = "right"
    model_input_names = ["input_ids", "attention_mask"]

    def get_vocab(self):
        return {self.eos_token: self.eos_token_id}

    def __call__(self, parts, add_special_tokens=False, **kw):
        return {"input_ids": [[(sum(map(ord, w)) * 31 + len(w)) % 1000 for w in p.split()] for p in parts]}


def test_processor_to_model_long_clip_two_chunks_device_mel_both_ways():
    """35 s clip -> 2 encoder chunks.  (a) processor computes the mel on the GPU (reference output contract, CUDA
    `audio_values`); (b) processor defers and the model computes mel + chunking from the raw waveform.  Same logits."""
    from oracle import model as om, processing as oproc
    from ultravox_b200.processing import MelSpec, UltravoxProcessor
    cfg, model, sd, sh = build()
    w = wave(3, 16000 * 35)
    text = "a b c <|audio|> d e"
    pa = UltravoxProcessor(MelSpec(feature_size=80), _Tok(), mel_device="cuda")
    ba = pa(text, audio=w, sampling_rate=16000)
    assert ba["audio_values"].is_cuda and tuple(ba["audio_values"].shape) == (2, 80, 3000)
    assert ba["audio_lens"].tolist() == [3000, 500] and ba["audio_token_len"].tolist() == [188, 32]
    ref = oproc.process(text, [w], lambda parts: _Tok()(parts)["input_ids"], 1000, n_mels=80)
    assert ba["input_ids"].tolist() == ref["input_ids"].tolist()
    assert np.abs(ba["audio_values"].cpu().numpy() - ref["audio_values"]).max() < 2e-3
    out_a = model(**{k: v for k, v in ba.items()})
    pb = UltravoxProcessor(MelSpec(feature_size=80), _Tok(), defer_mel=True)
    bb = pb(text, audio=w, sampling_rate=16000)
    assert "audio_values" not in bb and bb["audio_waveforms"].shape == (1, 16000 * 35)
    out_b = model(**{k: v for k, v in bb.items()})
    assert rel(out_b.logits, out_a.logits) < 2e-3
    ref_logits, _ = om.forward(sd, sh, ref_t(ref["input_ids"]), torch.from_numpy(ref["audio_values"]).to(torch.bfloat16).float(),
                               ref_t(ref["audio_token_start_idx"]), ref_t(ref["audio_lens"]), ref_t(ref["audio_token_len"]),
                               ref_t(ref["audio_batch_size"]))
    assert rel(out_a.logits, ref_logits) < 3e-2


def test_prefill_engine_long_clip_two_chunks():
    """35 s clip -> two encoder chunks inside the CUDA-graph prefill engine: same token and logits as the eager
    ``model.forward`` on the processor's output for the same waveform."""
    from ultravox_b200.engine import PrefillEngine
    from ultravox_b200.processing import MelSpec, UltravoxProcessor
    cfg, model, sd, sh = build()
    w = wave(3, 16000 * 35)
    pb = UltravoxProcessor(MelSpec(feature_size=80), _Tok(), defer_mel=True)
    bb = pb("a b c <|audio|> d e", audio=w, sampling_rate=16000)
    assert bb["audio_token_len"].tolist() == [188, 32]
    eng = PrefillEngine(model, 16000 * 35, bb["input_ids"], bb["audio_token_start_idx"], bb["audio_token_len"], bb["audio_batch_size"])
    assert eng.chunked and eng.graph is not None
    tok = eng.run_e2e(torch.from_numpy(w[None]).pin_memory()).clone()
    out = model(logits_to_keep=1, **{k: v for k, v in bb.items()})
    assert int(tok[0]) == int(out.logits.view(1, -1).argmax(-1))
    assert torch.allclose(eng.logits, out.logits.view(1, -1), rtol=1e-5, atol=1e-5)


def ref_t(a):
    return torch.from_numpy(np.asarray(a))


def test_cfg1_shapes_tiny_whisper_llama_1b_one_second_clip():
    """BASELINE config 1 shapes (Whisper-tiny + Llama-3.2-1B, 1 s clip): finite logits [1, 20, 128256], parity of the
    last-position logits with the fp32 oracle."""
    from oracle import logmel as ol, model as om
    from ultravox_b200 import ops
    cfg, model, sd, sh = build("tiny_1b")
    padded, batch = make_batch(cfg, [16000])
    assert batch["input_ids"].shape == (1, 20) and batch["audio_token_len"].tolist() == [7]
    mel = ops.logmel(torch.from_numpy(padded).cuda(), 80)
    assert tuple(mel.shape) == (1, 80, 100)
    out = model(audio_values=mel, **{k: v.cuda() for k, v in batch.items()})
    assert tuple(out.logits.shape) == (1, 20, 128256) and bool(torch.isfinite(out.logits).all())
    ref, _ = om.forward(sd, sh, batch["input_ids"], mel.cpu().to(torch.bfloat16).float(), batch["audio_token_start_idx"],
                        batch["audio_lens"], batch["audio_token_len"], batch["audio_batch_size"], last_only=True)
    got = out.logits[:, -1:].cpu()
    # random-init logits over a 128k vocabulary are nearly tied at the top: require the oracle's argmax in the GPU top-5
    assert rel(got, ref) < 3e-2 and int(ref.argmax(-1)) in got.view(-1).topk(5).indices.tolist()


def test_decode_engine_graph_matches_generate():
    """CUDA-graphed decode (device-side positions, GEMV linears) reproduces model.generate token for token."""
    from ultravox_b200 import ops
    from ultravox_b200.engine import DecodeEngine
    cfg, model, sd, sh = build()
    padded, batch = make_batch(cfg, [16000, 16000 + 500])
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    n_new = 6
    seq = model.generate(audio_values=mel, max_new_tokens=n_new, **{k: v.cuda() for k, v in batch.items()})
    S = batch["input_ids"].shape[1]
    emb = model._prepare_audio_embeds(batch["input_ids"].cuda(), mel, batch["audio_token_start_idx"], batch["audio_lens"],
                                      batch["audio_token_len"], batch["audio_batch_size"])
    for use_graph in (False, True):
        eng = DecodeEngine(model, 2, S + n_new + 2, use_graph=use_graph)
        toks = [eng.prefill(emb.clone()).clone()]
        for _ in range(n_new - 1):
            toks.append(eng.step().view(-1).clone())
        got = torch.stack(toks, 1)
        assert torch.equal(got, seq[:, S:]), (use_graph, got.tolist(), seq[:, S:].tolist())
    assert eng.launches_per_step > 20


def test_gemv_matches_gemm():
    from ultravox_b200 import ops
    for B, N, K in [(1, 512, 256), (3, 1024, 4096), (8, 256, 14336)]:
        g = torch.Generator().manual_seed(B)
        x = torch.randn(B, K, generator=g).bfloat16().cuda()
        w = (torch.randn(N, K, generator=g) * 0.05).bfloat16().cuda()
        r = torch.randn(B, N, generator=g).bfloat16().cuda()
        ref = x.float() @ w.float().T + r.float()
        assert rel(ops.gemv(x, w, residual=r), ref.to(torch.bfloat16)) < 1e-3
        assert rel(ops.gemv(x, w, out_dtype=torch.float32), x.float() @ w.float().T) < 1e-5


class _ChatTok(_Tok):
    """_Tok + what LocalInference needs: chat template, decode, left padding, eos-string splitting."""
    padding_side = "left"
    pad_token_id = 1000
    added_tokens_encoder = {"<|eot_id|>": 1000}

    def convert_tokens_to_ids(self, t):
        return self.added_tokens_encoder[t]

    def __call__(self, parts, add_special_tokens=False, **kw):
        out = []
        for p in parts:
            words = p.replace(self.eos_token, f" {self.eos_token} ").split()
            out.append([self.eos_token_id if w == self.eos_token else (sum(map(ord, w)) * 31 + len(w)) % 1000 for w in words])
        return {"input_ids": out}

    def apply_chat_template(self, messages, add_generation_prompt=True, tokenize=False, chat_template=None, **kw):
        text = " ".join(f"<s> {m['role']} : {m['content']} {self.eos_token}" for m in messages)
        return text + (" <s> assistant :" if add_generation_prompt else "")

    def decode(self, ids, skip_special_tokens=True):
        ids = [int(i) for i in (ids.tolist() if hasattr(ids, "tolist") else ids)]
        return " ".join(f"t{i}" for i in ids if not (skip_special_tokens and i == self.eos_token_id))

    def pad(self, features, padding=True, max_length=None, pad_to_multiple_of=None, return_tensors=None, **kw):
        import transformers
        L = max(len(f["input_ids"]) for f in features)
        out = {"input_ids": [], "attention_mask": []}
        for f in features:
            ids = list(map(int, f["input_ids"]))
            n = L - len(ids)
            out["input_ids"].append([self.pad_token_id] * n + ids)
            out["attention_mask"].append([0] * n + [1] * len(ids))
        out.update({k: [f[k] for f in features] for k in features[0] if k not in ("input_ids", "attention_mask")})
        return transformers.BatchFeature(out, tensor_type=return_tensors)


def test_local_inference_single_batch_stream_conversation():
    """the reference's LocalInference surface (ref inference/infer.py:20-342) over the B200 model:
    infer == infer_stream (same tokens, chunked), infer_batch (left-padded collate) agrees with per-sample infer,
    conversation mode re-uses the returned KV cache and prefills only the new turn, 48 kHz input is resampled."""
    from ultravox_b200.data_proc import VoiceSample
    from ultravox_b200.inference import InferenceChunk, InferenceStats, LocalInference
    from ultravox_b200.processing import MelSpec, UltravoxProcessor
    cfg, model, sd, sh = build()
    tok = _ChatTok()
    proc = UltravoxProcessor(MelSpec(feature_size=80), tok, mel_device="cuda")
    inf = LocalInference(model, proc, tok, conversation_mode=False)
    s1 = VoiceSample.from_prompt_and_raw("Listen to <|audio|> and answer", wave(1, 16000), 16000)
    s2 = VoiceSample.from_prompt_and_raw("<|audio|> what", (wave(2, 48000 * 2) * 3000).astype(np.int16), 48000)   # 2 s @ 48 kHz
    s3 = VoiceSample.from_prompt("plain text question without audio")
    o1 = inf.infer(s1, max_tokens=6)
    assert o1.output_tokens <= 6 and o1.input_tokens > 7 and o1.text.startswith("t")
    feats2 = inf._dataproc(s2)
    assert feats2["audio_lens"].tolist() == [200] and feats2["audio_token_len"].tolist() == [13]    # 2 s -> 200 frames -> 13 tokens
    # streaming: same tokens as infer, delivered as text deltas + final stats
    msgs = list(inf.infer_stream(s1, max_tokens=6))
    assert isinstance(msgs[-1], InferenceStats) and all(isinstance(m, InferenceChunk) for m in msgs[:-1])
    assert "".join(m.text for m in msgs[:-1]) == o1.text and msgs[-1].output_tokens == o1.output_tokens
    # batch (left padding; audio rows together, text-only rows together - the reference collator cannot mix them) vs one by one
    s4 = VoiceSample.from_prompt("another much longer plain text question to force left padding in the batch")
    first, rest = [], []
    for group in ([s1, s2], [s3, s4]):
        singles = [inf.infer(s, max_tokens=5) for s in group]
        batch = inf.infer_batch(group, max_tokens=5)
        assert len(batch) == 2 and all(b.input_tokens == max(s.input_tokens for s in singles) for b in batch)
        for sg, bt in zip(singles, batch):
            a, b = sg.text.split(), bt.text.split()
            first.append(a[0] == b[0])
            rest += [x == y for x, y in zip(a, b)]
    # random-init weights have near-tied logits, and a greedy sequence diverges for good after one flipped token: the strict
    # checks of the padded path are teacher-forced (test_left_padded_batch_forward_and_generate); here the first tokens
    # (pure prefill) must agree and the continuations mostly
    assert np.mean(first) >= 0.75 and np.mean(rest) >= 0.4, (first, np.mean(rest))
    sampled = inf.infer(s3, max_tokens=3, temperature=0.7)          # ref infer.py:319-328: temperature > 0 samples
    assert sampled.output_tokens >= 1 and sampled.input_tokens == inf.infer(s3, max_tokens=1).input_tokens
    # conversation mode: turn 2 only prefills its own suffix on top of the cache from turn 1
    conv = LocalInference(model, proc, tok, conversation_mode=True)
    c1 = conv.infer(s1, max_tokens=4)
    cache = conv.past_key_values
    assert cache is not None and cache.length == c1.input_tokens + c1.output_tokens - 1
    assert conv.past_messages[-1] == {"role": "assistant", "content": c1.text}
    assert conv.past_messages[0]["content"].count(tok.eos_token) == 7 and "<|audio|>" not in conv.past_messages[0]["content"]
    seen = []
    orig = model.llama_hidden
    model.llama_hidden = lambda emb, *a, **k: (seen.append(emb.shape[1]), orig(emb, *a, **k))[1]
    try:
        c2 = conv.infer(VoiceSample.from_prompt("and then what happened"), max_tokens=4)
    finally:
        model.llama_hidden = orig
    assert c2.input_tokens > c1.input_tokens + c1.output_tokens and seen[0] == c2.input_tokens - cache.length   # suffix only
    assert conv.past_key_values.length == c2.input_tokens + c2.output_tokens - 1 and len(conv.past_messages) == 4
    with pytest.raises(AssertionError):
        conv.infer_batch([s1])


def test_pipeline_end_to_end_and_repetition_penalty():
    """ref ultravox_pipeline.py: one call from raw audio to text; penalty 1.0 == plain greedy generate, the default 1.1 goes
    through the HF-exact logits processor (unit-tested on CPU) on the device logits of every step."""
    from ultravox_b200.data_proc import VoiceSample
    from ultravox_b200.inference import LocalInference
    from ultravox_b200.pipeline import UltravoxPipeline
    from ultravox_b200.processing import MelSpec, UltravoxProcessor
    cfg, model, sd, sh = build()
    tok = _ChatTok()
    proc = UltravoxProcessor(MelSpec(feature_size=80), tok, mel_device="cuda")
    pipe = UltravoxPipeline(model, tokenizer=tok, processor=proc)
    a = wave(4, 16000)
    plain = pipe({"audio": a, "sampling_rate": 16000, "prompt": "Listen to <|audio|> and answer"}, max_new_tokens=6, repetition_penalty=1.0)
    inf = LocalInference(model, proc, tok)
    ref = inf.infer(VoiceSample.from_prompt_and_raw("Listen to <|audio|> and answer", a, 16000), max_tokens=6)
    assert plain == ref.text
    pen = pipe({"audio": (a * 20000).astype(np.int16), "sampling_rate": 16000}, max_new_tokens=6)       # default penalty 1.1
    assert isinstance(pen, str) and 1 <= len(pen.split()) <= 6
    toks = pen.split()
    strong = pipe({"audio": a, "sampling_rate": 16000}, max_new_tokens=8, repetition_penalty=50.0).split()
    assert len(set(strong)) == len(strong)                  # a huge penalty never repeats a token it has already produced
    sampled = pipe({"audio": a, "sampling_rate": 16000}, temperature=0.7, max_new_tokens=4)      # temperature > 0 samples (ref infer.py:319-328)
    assert isinstance(sampled, str)


def test_llama_hidden_fused_paths_match_unfused():
    """Round-2 prefill path (pre-tiled weight images, RoPE in the q|k|v epilogue, SwiGLU in the gate|up epilogue) against the
    round-1 sequence of separate kernels on the row-major weights.  Cache and cache-less runs of one path are bit-identical;
    across the two paths the gate|up tiling (hence the fp32 summation order) and the sigmoid approximation differ, so hidden
    states agree to bf16 rounding noise, not bit for bit."""
    import ultravox_b200.model as mm
    from ultravox_b200.config import PRESETS, preset
    from ultravox_b200.model import UltravoxModel
    base = PRESETS["v0_5_8b"]
    cfg = preset("v0_5_8b", audio_config=dict(base["audio_config"], encoder_layers=1),
                 text_config=dict(base["text_config"], num_hidden_layers=2, vocab_size=2048))
    model = UltravoxModel(cfg, device="cuda").init_random_(seed=1)
    g = torch.Generator().manual_seed(0)
    emb = (torch.randn(1, 201, 4096, generator=g) * 0.5).to(torch.bfloat16).cuda()   # B=1, S=201: both paths pick the same tilings
    assert model._tiled_weights() is not None and model._tiled_weights()[0]["gate_up"].swiglu
    fused = model.llama_hidden(emb.clone()).clone()
    cache = model.new_cache(1, 210)
    fused_c = model.llama_hidden(emb.clone(), cache).clone()
    saved = (mm.USE_TILED, mm.FUSE_ROPE, mm.FUSE_SWIGLU)
    try:
        mm.USE_TILED = mm.FUSE_ROPE = mm.FUSE_SWIGLU = False
        model._tiled = None
        assert model._tiled_weights() is None
        plain = model.llama_hidden(emb.clone()).clone()
        cache2 = model.new_cache(1, 210)
        plain_c = model.llama_hidden(emb.clone(), cache2).clone()
    finally:
        mm.USE_TILED, mm.FUSE_ROPE, mm.FUSE_SWIGLU = saved
        model._tiled = None
    assert torch.equal(fused, fused_c) and torch.equal(plain, plain_c)
    assert rel(fused, plain) < 6e-3, rel(fused, plain)
    assert torch.equal(cache.k[0, :, :201], cache2.k[0, :, :201]) and torch.equal(cache.v[0, :, :201], cache2.v[0, :, :201])  # layer 0: same inputs
    assert rel(cache.k[1, :, :201], cache2.k[1, :, :201]) < 6e-3
    # opt-in weight-streaming GEMMs (UVX_GEMM_WS=1: tokens on the UMMA N dimension, stream-K, 128-row images incl. the pair-permuted
    # q|k|v image and the 16 | 16 gate|up image) on the same model: prefill at S = 201 and a one-token step on the cache
    from ultravox_b200 import _lib
    saved_ws = mm.USE_WS
    try:
        mm.USE_WS = True
        _lib.lib().uvx_debug_gemm_ws(1, 0, 0)
        model._tiled = None
        tw = model._tiled_weights()
        assert tw is not None and tw[0]["qkv"].rope_pairs and tw[0]["gate_up"].R == 128
        cache3 = model.new_cache(1, 210)
        ws = model.llama_hidden(emb.clone(), cache3).clone()
        step_in = emb[:, :1].clone()
        ws_step = model.llama_hidden(step_in.clone(), cache3).clone()
    finally:
        mm.USE_WS = saved_ws
        _lib.lib().uvx_debug_gemm_ws(-1, 0, 0)
        model._tiled = None
    plain_step = model.llama_hidden(step_in.clone(), cache2).clone()
    # every projection sums in another order and the RMSNorm reads the bf16-rounded residual stream (the reference's own order)
    # instead of the fp32 split-K sums: two more roundings per layer than between the two gemm_tc paths above
    assert rel(ws, plain) < 1.2e-2, rel(ws, plain)
    assert rel(ws_step, plain_step) < 1.5e-2, rel(ws_step, plain_step)
    assert rel(cache3.k[0, :, :201], cache2.k[0, :, :201]) < 4e-3         # layer 0 keys: same inputs, another fp32 summation order
