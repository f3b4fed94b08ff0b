"""ragroute_b200: B200-native retrieval hot path of sacs-epfl/ragroute.

Three drop-in surfaces over one C-ABI CUDA library (include/ragroute_b200.h):

* ``ragroute_b200.faiss_compat`` -- ``IndexFlatIP`` / ``read_index`` / ``normalize_L2``
  (replaces the ``faiss`` calls of ragroute/data_source.py)
* ``ragroute_b200.router.FusedCorpusRouter`` -- ``CorpusRoutingNN`` + fused select
  (ragroute/router.py)
* ``ragroute_b200.rerank`` -- ``rerank_medrag`` / ``rerank_wikipedia`` / ``merge_topk``
  (ragroute/rerank.py)

plus ``ragroute_b200.federation.Federation`` which chains them for a query batch.
Importing the package does not touch CUDA; the first object constructed does, and
fails loudly when no sm_100 device or no built library is present.
"""
__version__ = "0.1.0"
