"""CPU oracle for the GAN train/inference hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch (CPU, fp32) restatement of the reference's
algorithm for the hot path (generator, discriminator, losses, two-optimizer
step).  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product (``cross-modality-minipig-gan_b200/mpgan``) never imports it and
has no CPU fallback.

Pinning status: the reference ships no golden vectors, so the
oracle is pinned against *outputs of the reference's own classes run in the
authoring container* (``oracle/make_golden.py`` imports
``/root/reference/code/GAN/GAN_final.py`` and ``/root/reference/test_runs/GAN.py``
verbatim through ``oracle/ref_shim.py``) and the resulting fixtures are
committed under ``tests/golden/``.  The one part that cannot be pinned is the
third-party ``monai==0.4.0`` ``UNet`` (not vendored by the reference, not
installable here): it is restated from its published source in
``oracle/monai_unet.py`` -- for that module alone parity is "unpinned".
"""
