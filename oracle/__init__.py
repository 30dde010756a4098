"""TEST INFRASTRUCTURE ONLY — CPU/torch-fp32 restatements of the reference's hot path.

Nothing under ``oracle/`` is part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it, and only as the checker or
the timed CPU baseline — never as a fallback for the CUDA library.

Pinning status:
  * UNet forward/backward + WeightedCrossEntropyLoss: the reference holds no value tests (only a
    shape assert, models/unet_model.py:219-222). The restatement in ``unet_ref.py`` is pinned
    against the reference ITSELF, imported from /root/reference in the authoring container
    (tests/test_oracle.py) and through golden vectors generated from it
    (oracle/make_golden.py -> tests/golden/unet_golden.npz).
  * get_instance_masks: pinned by 84 shipped mask -> instance pairs of the reference
    (data/raw/processed/predictions/DIC-C2DH-HeLa/01_RES{,_INST}); a subset is committed as
    tests/golden/ccl_golden.npz.
  * overlap-tile inference: absent from the reference -> PARITY UNPINNED; semantics
    are defined in ``overlap_tile_ref.py`` and checked through the aligned-tile invariant
    (tests/test_overlap_tile_cpu.py on the oracle itself, tests/test_tiling_gpu.py on the library).
  * calculate_weight_map (row N4): ``weight_map_ref.py`` is pinned by the 84 float64 weight maps the
    reference stores next to its instance masks (data/raw/train/DIC-C2DH-HeLa/01_ST/WEIGHT_MAPS; all
    checked live, three committed in tests/golden/weight_map_golden.npz) and by the live function.
"""
