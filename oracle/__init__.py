"""
oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU restatement (numpy / scipy / plain-Python loops) of the one HyPRec hot path this
repository re-implements on B200: the confidence-weighted ALS fit
(/root/reference/lib/collaborative_filtering.py:69-99, 224-259), the dense prediction
(:262-294) and the top-k ranking / evaluation report that
/root/reference/lib/evaluator.py:215-341, /root/reference/lib/abstract_recommender.py:100-187
and /root/reference/util/top_recommendations.py:23-40 run over it.

Who may import this package: tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py -- always as the checker or the timed CPU baseline, never
as the thing shipped.  Nothing under hyprec_b200/ imports it; the product path raises when
the CUDA library is missing instead of falling back to this code.

Pinning: the reference's own tests hold no golden factor values for this path
 -- only MRR / nDCG known answers (tests/metrics_tests.py:60-107) and
container invariants.  The oracle is therefore pinned (a) against those known answers and
(b) against outputs of the UNMODIFIED reference classes executed in the build container
(oracle/reference_loader.py imports them from /root/reference; tests/golden/make_golden.py
writes the vectors that tests/golden/*.npz hold).  /root/reference does not exist on the GPU
box: only the committed vectors travel.
"""
