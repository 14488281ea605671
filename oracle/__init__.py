"""TEST INFRASTRUCTURE ONLY.

CPU restatements (the *oracle*) of the reference hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; the product package ``mtg_card_image_segmentation_b200``
never does (tests/test_layout.py enforces that).
"""
