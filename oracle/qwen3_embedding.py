"""Qwen3-Embedding / Qwen3-Reranker -- restates /root/reference/src/models/qwen3_embedding/mod.rs:38-64,
src/models/qwen3_reranker/mod.rs:23-31 and l2_normalize / cosine_similarity_no_l2
(src/models/common/modules.rs:1287-1294,1381-1389).  Token ids in (the tokenizer is out of scope)."""
import numpy as np

from .qwen3 import Qwen3Model

F32 = np.float32


def l2_normalize(t, axis=-1):
    """t / sqrt(sum(t^2) + 1e-6)"""
    t = t.astype(F32)
    return (t / np.sqrt(np.sum(t * t, axis=axis, keepdims=True, dtype=F32) + F32(1e-6))).astype(F32)


class Qwen3Embedding:
    def __init__(self, cfg, w):
        self.model = Qwen3Model(cfg, w, [])

    def embed_one(self, ids):
        """forward_hidden(ids, offset 0) -> last-token hidden after the final RMSNorm -> L2 normalise; cache cleared."""
        hidden = self.model.forward_hidden(np.asarray(ids).reshape(1, -1), None, 0)[0]   # (1, H)
        self.model.clear_cache()
        return l2_normalize(hidden, -1)[0]

    def embed_multi(self, list_of_ids):
        if len(list_of_ids) == 0:
            raise ValueError("embedding input cannot be empty")
        return np.stack([self.embed_one(i) for i in list_of_ids], 0)


class Qwen3Reranker:
    def __init__(self, cfg, w):
        self.embedding = Qwen3Embedding(cfg, w)

    def rerank(self, query_ids, documents_ids):
        q = self.embedding.embed_one(query_ids)[None]
        d = self.embedding.embed_multi(documents_ids)
        return np.matmul(q, d.T)[0].astype(F32)     # cosine_similarity_no_l2: the embeddings are already unit vectors
