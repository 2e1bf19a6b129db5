"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference algorithms on the hot path
(nazimurahman/humanoid-vision-system: src/models/manifold_layers.py,
src/models/yolo_head.py, src/inference/postprocessing.py).

Nothing under this package is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it -- as the checker or the timed CPU
baseline, never as a fallback for the CUDA path.

Parity pinning: the restatement is checked, in this container, against the
reference's own Python modules (imported from /root/reference with the
repair set R1/R5/R7/R9 of SURVEY.md Appendix A, see
``oracle/reference_repaired.py``) by ``oracle/make_golden.py``; that script
also writes the golden vectors under ``tests/golden/`` that travel to the GPU
box.  The one known-answer vector the reference's tests hold for this path
(src/tests/test_inference.py:361-379, the 3-box NMS case) is part of the
golden set.
"""
