"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy/scipy, fp64) of the zoom-FFT PSD hot path of
alfille/pypanadapter.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package; the product (``pypanadapter_b200``) never does and has no CPU
fallback.

Pinning status (see DESIGN.md "Oracle"):
* The reference ships NO tests, golden vectors or fixtures, so
  the oracle is pinned against *outputs of the reference's own lines run in
  the build container*: ``oracle/ref_harness.py`` exec-loads the unmodified
  ``/root/reference/pypanadapter_{spectrum,thread}.py`` under Qt/pyqtgraph
  stubs and calls ``ApplicationDisplay.update/zoomfft``, ``PSD.update``,
  ``Data.add`` and ``Waterfall.image_update`` themselves;
  ``oracle/make_golden.py`` records those outputs as ``tests/golden/*.npz``
  and ``tests/test_oracle.py`` checks the restatement against them.
* ``pyrtlsdr``'s uint8 -> complex conversion is an un-vendored dependency that
  is not installed here: that single function is restated from its published
  algorithm and is "parity unpinned".
"""
