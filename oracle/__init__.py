"""TEST INFRASTRUCTURE ONLY — the parity oracle for the JPerceiver training hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker or the timed CPU baseline.  ``jperceiver_b200`` never imports this package.

Contents
--------
``port.py``        CPU fp32 (torch) functional restatement of the reference's hot path
                   (network forward + ``compute_losses``), every function citing the
                   reference file:line it follows.  This is what travels to the GPU box.
``ref_loader.py``  Imports the reference's *own* modules from ``/root/reference`` under an
                   import shim (authoring container only; the GPU box has no reference).
``make_golden.py`` Runs the real reference and writes ``tests/golden/*.npz`` fixtures (KAT vectors, one full
                   forward+backward run per ``opt.type``).
``make_golden_bev.py`` / ``make_golden_eval.py`` / ``make_golden_pipeline.py`` / ``make_golden_sampler.py``
                   The same for the BEV loss variants, the validation metrics (``pixel_error.py``),
                   ``MonoDataset.preprocess`` and the samplers' index sequences.
``eval_port.py`` / ``pipeline_port.py``  Restatements of the validation-hook body and of ``preprocess``.

Parity status: the reference ships no tests or golden vectors.  The port is pinned against
(a) outputs of the reference itself executed here through ``ref_loader`` (fixtures committed
under ``tests/golden`` with the generating script) and (b) the closed-form known-answer
vectors of SURVEY.md §8c.  Third-party arithmetic that is absent from ``/root/reference``
(torchgeometry ``warp_perspective``, skimage ``find_boundaries``) is restated from the
published algorithm — those two pieces are "parity unpinned" and say so where they are used.
"""
