"""CPU oracle for the UDA segmentation hot path — TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a plain-PyTorch / numpy / C restatement of the
reference algorithms (bempt/uda_aerial_semantic_segmentation_research) used as
the *checker* for the CUDA kernels.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product package (``uda_aerial_semantic_segmentation_research_b200``) never
imports from here and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * losses / discriminator / metrics restatements are pinned against the
    reference's own modules imported from /root/reference (``gen_golden.py``
    writes ``tests/golden/*.npz``; ``tests/test_oracle.py`` re-checks live when
    the reference tree is present) and against the survey's closed-form KATs.
  * the U-Net restatement (``ref_unet.py``) follows the third-party
    ``segmentation_models_pytorch`` (>=0.3.0, unpinned, NOT vendored in the
    reference) + torchvision ResNet; the reference tests hold no numeric pin
    for it  ->  **parity unpinned** for U-Net values beyond structure/shape.
"""
