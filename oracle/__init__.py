"""CPU oracle for the HermNet message-passing hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  Nothing under ``hermnet_b200/`` imports
``oracle``; the product path raises when the CUDA library is missing.

Contents
--------
* ``hermnet_oracle``  -- pure-PyTorch (CPU, fp32 or fp64) restatement of the reference
  forward pass (``/root/reference/HermNet/hermnet.py:118-152``, ``rmnet.py:11-193``,
  ``utils.py:11-24,138-160``) written against the reference's ``state_dict`` layout,
  plus the builder-owned HPNet / HTNet specification.
* ``neighbor_oracle`` -- numpy brute-force restatement of the neighbour-list contract of
  ``HermNet/data.py:14-24`` (ASE ``primitive_neighbor_list`` / torch_cluster
  ``radius_graph`` semantics) and the canonical triplet list.
* ``nl_oracle.c``     -- the same neighbour-list contract in plain C for sizes numpy
  cannot brute-force in seconds (built by ``oracle/build.py`` into ``oracle/_build``).
* ``ref_shims``       -- minimal stand-ins for the third-party modules the reference
  imports (torch_geometric, torch_scatter, torch_cluster, ase) so that the reference's
  OWN source files can be executed in the authoring container to pin this oracle
  (see ``tests/golden/make_golden.py``).

Parity pinning
--------------
The reference ships no tests, golden vectors or fixtures and its
third-party dependencies are not installable here.  The oracle is pinned by executing
the reference's own ``HermNet/hermnet.py`` / ``rmnet.py`` / ``utils.py`` (unmodified,
imported from ``/root/reference``) over ``ref_shims`` and comparing; the committed
fixtures under ``tests/golden`` were produced by that run.  The semantics of the shimmed
third-party calls themselves (PyG ``propagate``, ``GaussianSmearing``, ``scatter``,
ASE's neighbour list) remain "[upstream, unverified here]", and HPNet/HTNet do not exist
in the reference at all: for those two **parity is unpinned** (builder-owned spec).
"""
