"""Sampling -- restates /root/reference/src/models/common/sample.rs:7-60 (`get_logit_processor`, `use_repeat_penalty`) and
the third-party pieces it calls, which are NOT under /root/reference:

  * candle-transformers 0.9.2 `generation::LogitsProcessor` (Sampling::{ArgMax, All, TopP, TopK, TopKThenTopP}):
    prs = softmax(logits / temperature); top-p = sort by descending probability, zero every entry once the running sum
    of the kept ones has reached p; top-k = keep the k largest; then `rand::distr::weighted::WeightedIndex::new(prs)` and
    one draw from `StdRng::seed_from_u64(seed)`;
  * candle-transformers `utils::apply_repeat_penalty`: every DISTINCT token of the context, logit >= 0 ? / p : * p;
  * rand 0.9 / rand_chacha 0.9 / rand_core 0.9: StdRng = ChaCha12 (64-bit block counter, stream 0), the 32-byte seed filled
    by PCG32 from the u64 seed, one u32 word per draw, Uniform<f32>(0, total) = ((word >> 9 | exp 0) - 1.0) * total,
    WeightedIndex = partition_point(cumulative <= draw).

**Unverifiable here** (no Rust toolchain, crates not vendored): the restatement follows the published algorithms; the ChaCha
core is pinned by the RFC 7539 block test vector (20 rounds; StdRng runs the same core with 12).  Two places where a
bit-for-bit match with the crates is not attempted, by design, because they are not properties of the algorithm:
  (1) summation ORDER of fp32 sums over the vocabulary (softmax denominator, cumulative weights): candle / rand add left to
      right; here -- and in the CUDA sampler, which this file is the oracle of -- sums run over 256-element chunks left to
      right and then over the chunk totals left to right (`blocked_*`).  Same values up to fp32 rounding of the bounds;
  (2) the ORDER of the k candidates of top-k sampling: candle takes whatever `select_nth_unstable_by` leaves in the first k
      slots (an implementation detail of Rust's pdqselect); here the candidates are sorted by descending probability, ties
      by ascending token id.  Same distribution, possibly a different token for the same seed.
This module is test infrastructure (see oracle/__init__.py)."""
import numpy as np

F32 = np.float32
CHUNK = 256
MASK32 = 0xFFFFFFFF


# ----------------------------------------------------------------------------- StdRng = ChaCha12, seeded through PCG32
def _rotl(x, n):
    return ((x << n) | (x >> (32 - n))) & MASK32


def chacha_block(key_words, counter, rounds, nonce_words=(0, 0)):
    """One 64-byte ChaCha block as 16 u32 words.  Layout of rand_chacha (and djb's original): words 12-13 = 64-bit
    block counter, words 14-15 = stream id."""
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + \
         [counter & MASK32, (counter >> 32) & MASK32, nonce_words[0], nonce_words[1]]
    x = list(st)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & MASK32; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & MASK32; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & MASK32; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & MASK32; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & MASK32 for a, b in zip(x, st)]


def seed_from_u64(state):
    """rand_core::SeedableRng::seed_from_u64: the 32-byte seed is eight PCG32 outputs (little endian)."""
    MUL, INC = 6364136223846793005, 11634580027462260723
    words = []
    for _ in range(8):
        state = (state * MUL + INC) & 0xFFFFFFFFFFFFFFFF
        xorshifted = (((state >> 18) ^ state) >> 27) & MASK32
        rot = state >> 59
        words.append(((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & MASK32)
    return words


class StdRng:
    """Stateless view of the stream: word(n) is the n-th u32 that `StdRng::seed_from_u64(seed)` returns."""

    def __init__(self, seed):
        self.key = seed_from_u64(int(seed) & 0xFFFFFFFFFFFFFFFF)

    def word(self, n):
        return chacha_block(self.key, n // 16, 12)[n % 16]

    def uniform01(self, n):
        """Uniform<f32>: 23 random mantissa bits with exponent 0, minus one."""
        bits = np.uint32((self.word(n) >> 9) | 0x3F800000)
        return F32(bits.view(F32) - F32(1.0))


# ----------------------------------------------------------------------------- blocked fp32 arithmetic over the vocabulary
def blocked_sum(x):
    """sum over 256-element chunks left to right, then over the chunk totals left to right (all fp32)."""
    x = np.asarray(x, F32)
    n = (len(x) + CHUNK - 1) // CHUNK
    pad = np.zeros(n * CHUNK, F32); pad[:len(x)] = x
    tot = np.cumsum(pad.reshape(n, CHUNK), axis=1, dtype=F32)[:, -1]
    return np.cumsum(tot, dtype=F32)[-1]


def blocked_pick(w, draw01):
    """WeightedIndex over weights w (fp32, >= 0) with blocked cumulative sums: cumulative[i] = prefix of the chunk totals
    + running sum inside the chunk; the draw is `draw01 * total`; returns the number of cumulative values <= draw."""
    w = np.asarray(w, F32)
    V = len(w)
    n = (V + CHUNK - 1) // CHUNK
    pad = np.zeros(n * CHUNK, F32); pad[:V] = w
    within = np.cumsum(pad.reshape(n, CHUNK), axis=1, dtype=F32)
    tot = within[:, -1]
    pre = np.zeros(n, F32)
    pre[1:] = np.cumsum(tot, dtype=F32)[:-1]
    total = F32(pre[-1] + tot[-1])
    if not (total > 0) or not np.isfinite(total):
        raise ValueError("WeightedIndex: invalid weights")
    u = F32(draw01 * total)
    cum = (pre[:, None] + within).astype(F32).reshape(-1)[:V]
    return int(min(np.count_nonzero(cum <= u), V - 1))


# ----------------------------------------------------------------------------- the reference's call sites
def apply_repeat_penalty(logits, penalty, context):
    logits = np.array(logits, F32)
    seen = set()
    for t in context:
        t = int(t)
        if t in seen:
            continue
        seen.add(t)
        if t < len(logits):
            logits[t] = logits[t] / F32(penalty) if logits[t] >= 0 else logits[t] * F32(penalty)
    return logits


def use_repeat_penalty(repeat_penalty, repeat_last_n, logits, context):
    """sample.rs:40-60."""
    if repeat_penalty == 1.0 or repeat_last_n == 0:
        return np.asarray(logits, F32)
    start = max(0, len(context) - repeat_last_n) if repeat_last_n is not None else 0
    return apply_repeat_penalty(logits, repeat_penalty, context[start:])


def sampling_mode(temperature, top_p, top_k):
    """get_logit_processor (sample.rs:7-38) + LogitsProcessor::new: -> ('argmax' | 'all' | 'topp' | 'topk' | 'topk_topp')."""
    temperature = None if (temperature is None or temperature < 1e-7) else temperature
    if temperature is None:
        return "argmax"
    if top_k is None:
        return "all" if top_p is None else "topp"
    return "topk" if top_p is None else "topk_topp"


def softmax_blocked(logits, temperature):
    """prs = softmax_last_dim(logits / temperature): the tensor is multiplied by (1 / temperature) rounded to f32 (candle's
    `Tensor / f64` is an affine with mul = 1/t)."""
    x = (np.asarray(logits, F32) * F32(1.0 / float(temperature))).astype(F32)
    e = np.exp((x - x.max()).astype(F32)).astype(F32)
    return (e / blocked_sum(e)).astype(F32)


def _keys(p):
    return np.asarray(p, F32).view(np.uint32)        # p >= 0: the bit pattern orders like the value


def topp_weights(p, top_p):
    """sample_topp: keep, in order of descending probability (ties: ascending id), entries while the sum of the kept ones is
    < top_p.  Stated without a sort: T = smallest key with A(T) := blocked_sum(p[key > T]) < top_p; everything above T is
    kept; the j-th (by id) entry equal to T is kept iff A(T) + j * value < top_p."""
    p = np.asarray(p, F32)
    k = _keys(p)
    lo, hi = 0, 0x7F800000
    while lo < hi:                                   # 31 halvings
        mid = (lo + hi) // 2
        if blocked_sum(np.where(k > mid, p, F32(0))) < F32(top_p):
            hi = mid
        else:
            lo = mid + 1
    T = lo
    A = blocked_sum(np.where(k > T, p, F32(0)))
    keep = k > T
    ties = np.nonzero(k == T)[0]
    if len(ties):
        v = p[ties[0]]
        j = np.arange(len(ties), dtype=F32)
        keep[ties[(A + j * v).astype(F32) < F32(top_p)]] = True
    return np.where(keep, p, F32(0)).astype(F32)


def topk_candidates(p, k):
    """the k largest probabilities, sorted by descending value, ties by ascending id -> (values, ids)."""
    p = np.asarray(p, F32)
    order = np.lexsort((np.arange(len(p)), -p.astype(np.float64)))[:k]
    return p[order].copy(), order


def _pick_sequential(w, draw01):
    """WeightedIndex over a short list: cumulative sums left to right."""
    w = np.asarray(w, F32)
    cum = np.cumsum(w, dtype=F32)
    total = cum[-1]
    if not (total > 0) or not np.isfinite(total):
        raise ValueError("WeightedIndex: invalid weights")
    return int(min(np.count_nonzero(cum <= F32(draw01 * total)), len(w) - 1))


def sample(logits, temperature, top_p, top_k, rng, draw_index):
    """LogitsProcessor::sample for the processor `get_logit_processor(temperature, top_p, top_k, seed)` would build;
    `draw_index` = how many tokens this processor has sampled before (ArgMax consumes no randomness)."""
    mode = sampling_mode(temperature, top_p, top_k)
    logits = np.asarray(logits, F32).reshape(-1)
    if mode == "argmax":
        return int(np.argmax(logits))
    p = softmax_blocked(logits, temperature)
    u = rng.uniform01(draw_index)
    if mode == "all" or (mode == "topp" and (top_p <= 0.0 or top_p >= 1.0)):
        return blocked_pick(p, u)
    if mode == "topp":
        return blocked_pick(topp_weights(p, top_p), u)
    if top_k >= len(p):
        return blocked_pick(p, u) if mode == "topk" else blocked_pick(topp_weights(p, top_p), u)
    vals, ids = topk_candidates(p, top_k)
    if mode == "topk_topp":
        sum_p = np.cumsum(vals, dtype=F32)[-1]
        if not (top_p <= 0.0 or top_p >= sum_p):
            cum = F32(0)
            for j in range(len(vals)):               # sample_topp on the candidate list (already in descending order)
                if cum >= F32(top_p):
                    vals[j] = F32(0)
                else:
                    cum = F32(cum + vals[j])
    return int(ids[_pick_sequential(vals, u)])


class Sampler:
    """GenerationContext's sampling state (generate.rs:21-86): processor + repeat penalty; sample_and_push order."""

    def __init__(self, temperature=None, top_p=None, top_k=None, repeat_penalty=None, repeat_last_n=None, seed=299792458):
        self.temperature, self.top_p, self.top_k = temperature, top_p, top_k
        self.repeat_penalty = 1.0 if repeat_penalty is None else repeat_penalty
        self.repeat_last_n = 64 if repeat_last_n is None else repeat_last_n
        self.rng = StdRng(seed)
        self.draws = 0

    def sample_and_push(self, logits, generated, penalise=True):
        logits = np.asarray(logits, F32).reshape(-1)
        if penalise:
            logits = use_repeat_penalty(self.repeat_penalty, self.repeat_last_n, logits, generated)
        tok = sample(logits, self.temperature, self.top_p, self.top_k, self.rng, self.draws)
        if sampling_mode(self.temperature, self.top_p, self.top_k) != "argmax":
            self.draws += 1
        generated.append(tok)
        return tok
