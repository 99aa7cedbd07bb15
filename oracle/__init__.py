"""CPU oracle for the gradient-weighted attention-relevancy path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
it, and there only as the checker (or as the timed CPU baseline), never as the thing shipped.  The
product path (``transformer-mm-explainability_b200``) never imports this package and fails loudly when
its CUDA library is missing.

Parity status: the reference ships NO tests / golden vectors for the relevancy path, so
the oracle is pinned against *outputs of the reference itself run in the build container*:
``oracle/make_golden.py`` imports the unmodified reference from ``/root/reference`` (with the import shims
of ``oracle/ref_shims.py``), runs it on seeded inputs and commits the results under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks this restatement against those fixtures everywhere, and
``tests/test_oracle_vs_reference.py`` re-checks it against the live reference whenever ``/root/reference``
is present.  The ViT-B/16 *model* forward is the one exception ("parity unpinned": its source is not
vendored in the reference tree); its rule is pinned through ``avg_heads`` /
``apply_self_attention_rules``.
"""
