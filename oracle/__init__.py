"""CPU oracle for the DDSP-Piano synthesis hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
or as the timed CPU baseline -- never as a fallback for the CUDA path.

Parity status (see DESIGN.md "Oracle"):
  * the reference's own Python for this path (``ddsp_piano/modules/inharm_synth.py``,
    ``filtered_noise_synth.py:12-42``, ``polyphonic_dag.py``) was EXECUTED in the
    build container on top of a NumPy stand-in for ``tensorflow``/``gin``/``ddsp``
    (``oracle/tf_shim``) to generate ``tests/golden/*.npz``; the restatement in
    ``oracle/ddsp_piano_np.py`` is checked against those vectors.
  * the arithmetic underneath (``ddsp.core`` v3.7.0 and the TensorFlow CPU kernels
    it lowers to) is NOT vendored in the reference and is not installable here, so
    ``oracle/ddsp_core_np.py`` restates its published algorithm.  That layer is
    "parity unpinned": it is pinned only by the analytical known-answer tests of
    SURVEY.md section 8c, not by outputs of TensorFlow itself.
"""
