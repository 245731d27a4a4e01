"""CPU oracle for the point->BEV front end.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and there only as the checker or the timed CPU baseline.
The product package (``practical-collab-perception_b200`` / alias ``pcp_b200``)
never imports this package and raises if its CUDA library is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this
path.  The oracle is therefore pinned against OUTPUTS OF THE
REFERENCE ITSELF, produced in the build container by importing the reference's
own modules from ``/root/reference`` (``oracle/ref_loader.py``) and committed as
fixtures under ``tests/golden/`` by ``oracle/gen_golden.py``.
"""
