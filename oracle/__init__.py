"""CPU oracle for the NNUEEHCS uncertainty-estimation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``nnueehcs_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or as the timed CPU baseline -- never as the thing shipped.

Contents
--------
``uq_oracle``       torch-CPU restatement of ``EnsembleModel.forward``,
                    ``MCDropoutModel.forward`` (reference ``nnueehcs/models.py:99-163``)
                    and of the third-party ``deltaUQ_MLP.forward`` that
                    ``DeltaUQMLP.forward`` (``models.py:313-341``) delegates to.
``metrics_oracle``  numpy restatement of ``scipy.stats.wasserstein_distance`` and of
                    ``JensenShannonEvaluation.pdf_jsd`` (``nnueehcs/evaluation.py:175-188,
                    268-276``).
``shims``           import shims that let the *unmodified* reference package be imported
                    from ``/root/reference`` in the authoring container (used only by
                    ``tests/golden/make_golden.py`` to pin the oracle; cannot travel).

Parity pins (see DESIGN.md "Oracle"):
  * ensemble / MC-dropout / metrics : pinned against outputs of the reference's own
    code (run through ``shims``) and scipy, committed under ``tests/golden/``.
  * Delta-UQ : **parity unpinned** -- the arithmetic lives in the un-vendored,
    un-pinned ``deltauq`` package that is absent from the reference tree.
"""
