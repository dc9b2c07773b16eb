"""retrieval-fuse_b200: the RetrievalFuse hot path (patch encoders -> exact kNN
over the embedding bank -> compose -> U-Net + patch-attention refinement) as
hand-written sm_100a CUDA behind the reference's own module API.

Layout
  csrc/      CUDA kernels + the C ABI (include/rf_b200.h) -> librf_b200.so
  _lib.py    ctypes binding (fails loudly when the .so is missing)
  ops.py     tensor-level wrappers
  model/     drop-in `model` package of the reference (same classes, same
             state_dict keys): retrieval.py, attention.py, unet.py, refinement.py
  util/      drop-in `util.patcher`, `util.retrieval`
  dataset/   SceneHandler-compatible chunk access for in-memory scenes
  sharded.py bank sharding across ranks + all-gather/merge of per-shard top-k
"""
__version__ = "0.1.0"
