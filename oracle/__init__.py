"""CPU oracle for the two SOFIMA hot paths (TEST INFRASTRUCTURE ONLY).

This package restates, in plain NumPy (fp32) and plain C, the algorithms of

  * /root/reference/mesh.py        (inplane_force, elastic_mesh_3d,
                                    velocity_verlet, relax_mesh)
  * /root/reference/flow_field.py  (masked_xcorr, _batched_xcorr,
                                    _batched_peaks, _peak_stats,
                                    JAXMaskedXCorrWithStatsCalculator.flow_field)

It is the checker for the CUDA path, never the product.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  Nothing under `sofima_b200/` imports it; the product
fails loudly when the CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * pinned against every known-answer test the reference holds for these paths
    (tests/flow_field_test.py, tests/mesh_test.py), ported in
    tests/test_oracle_kat.py;
  * pinned against golden vectors produced by running the reference's OWN
    source files (mesh.py, flow_field.py from /root/reference) on a NumPy
    stand-in for the JAX primitives (tests/golden/make_golden.py; JAX itself is
    not installable in this image, so the semantics of jax.lax.dynamic_slice /
    conv_general_dilated_patches / .at[].set / argmax are restated from their
    documentation, not executed).
"""
