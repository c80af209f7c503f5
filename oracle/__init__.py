"""CPU oracle of the TriCoLo embedding-similarity hot path — TEST INFRASTRUCTURE ONLY.

A plain NumPy / CPU-PyTorch restatement of
  * tricolo/loss/nt_xent.py            (NT-Xent / InfoNCE loss and its gradient)
  * tricolo/evaluation/eval_retrieval.py (text->shape retrieval + RR@k / NDCG@k / MRR)
used as the checker by tests/, by __graft_entry__.smoke() and as the timed CPU
baseline of bench.py (cpu_baseline / --impl reference).  Nothing in the product
package tricolo_b200/ imports it.

Pinning: the reference ships no tests or golden vectors.  The
oracle is pinned against outputs of the reference code itself, run unmodified in
the build container through a lightning/jsonlines shim by
tests/golden/make_golden.py; the resulting vectors are committed under
tests/golden/ and checked by tests/test_oracle_golden.py.  Tie order is
undefined in the reference (np.argsort introsort); on tied inputs the oracle is
the stated restatement (similarity descending, gallery index ascending).
"""
