"""Autoregressive loop -- restates /root/reference/src/models/common/generate.rs:21-159 and
sample.rs:7-60 for the deterministic (ArgMax) sampler.  Non-greedy sampling depends on candle's
LogitsProcessor RNG (StdRng::seed_from_u64) which is not restated; requesting it raises."""
import time
import numpy as np


class GenerationContext:
    """generate.rs:21-68."""

    def __init__(self, temperature=None, top_p=None, top_k=None, repeat_penalty=None, repeat_last_n=None,
                 seed=299792458, initial_seq_len=0, max_tokens=1024):
        temperature = None if (temperature is None or temperature < 1e-7) else temperature  # sample.rs:13
        if temperature is not None:
            raise NotImplementedError("oracle restates Sampling::ArgMax only")
        self.repeat_penalty = 1.0 if repeat_penalty is None else repeat_penalty
        self.repeat_last_n = 64 if repeat_last_n is None else repeat_last_n
        self.seqlen_offset = 0
        self.seq_len = initial_seq_len
        self.sample_len = max_tokens

    def prepare_for_next_token(self, token):
        self.seqlen_offset += self.seq_len
        self.seq_len = 1
        return np.array([[token]], dtype=np.uint32)


def apply_repeat_penalty(logits, penalty, context):
    """candle_transformers::utils::apply_repeat_penalty: for each distinct token in context,
    logit >= 0 ? logit / p : logit * p."""
    logits = logits.copy()
    for t in set(int(c) for c in context):
        if t < logits.shape[0]:
            logits[t] = logits[t] / penalty if logits[t] >= 0 else logits[t] * penalty
    return logits


def sample_and_push(ctx, logits, generated):
    """generate.rs:70-86."""
    logits = np.asarray(logits, dtype=np.float32).reshape(-1)
    if not (ctx.repeat_penalty == 1.0 or ctx.repeat_last_n == 0):  # sample.rs:40-60
        start = max(0, len(generated) - ctx.repeat_last_n)
        logits = apply_repeat_penalty(logits, ctx.repeat_penalty, generated[start:])
    token = int(np.argmax(logits))  # first maximal index
    generated.append(token)
    return token


def generate_generic(model, input_ids, data, ctx):
    """generate.rs:115-159 minus tokenizer/response building.
    Returns (generated ids, prompt_secs, completion_secs).  The first token is never
    EOS-checked; an EOS token is pushed before the break; cache cleared at the end."""
    generated = []
    eos = model.stop_token_ids()
    t0 = time.perf_counter()
    logits = model.forward_initial(input_ids, ctx.seqlen_offset, data)
    tok = sample_and_push(ctx, logits, generated)
    prompt_secs = time.perf_counter() - t0
    ids = ctx.prepare_for_next_token(tok)
    t0 = time.perf_counter()
    for _ in range(1, ctx.sample_len):
        logits = model.forward_step(ids, ctx.seqlen_offset)
        tok = sample_and_push(ctx, logits, generated)
        if tok in eos:
            break
        ids = ctx.prepare_for_next_token(tok)
    completion_secs = time.perf_counter() - t0
    model.clear_cache()
    return generated, prompt_secs, completion_secs
