"""CPU oracle for the SPGNN GNN stage — TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker or
the timed CPU baseline.  Nothing under ``spgnn_b200/`` imports it; the product
path fails loudly if its CUDA library is missing instead of falling back here.

What it restates
----------------
The reference (/root/reference, 100 % Python) keeps the hot path's arithmetic
in an un-vendored, un-pinned dependency: DGL (``models.py:8``; built from git
``master`` in ``docker_base/Dockerfile:130``; README floor 0.6.x; inferred
0.7.x).  DGL is not installable here, so ``oracle.dgl_ops`` restates the
published DGL-0.7.x algorithms (GATConv, GraphConv, SAGEConv-pool,
GINConv-mean, edge_softmax, ``dgl.batch``, graph construction) as the same op
sequence DGL issues, ``oracle.models`` restates the wiring of
``models.py:160-194, 283-340, 343-400, 403-540, 650-696`` and ``oracle.pe``
restates ``job_runner.py:1684-1702`` and ``:1712-1777``.

Parity status: PARTIALLY PINNED.  The reference ships no tests or golden
vectors.  What *is* pinned, by executing the reference's own
code in the authoring container (``tests/golden/make_golden.py``):
  * model wiring — ``/root/reference/models.py`` classes imported with
    ``oracle.dgl_ops`` injected as ``dgl.nn.pytorch`` → ``tests/golden/wiring_*.npz``;
  * graph construction edge order — the literal networkx calls of
    ``job_runner.py:1779-1801`` → ``tests/golden/graph_*.npz``;
  * positional encodings and anchor selection — function bodies of
    ``job_runner.py:1684-1777`` executed through a minimal graph shim →
    ``tests/golden/pe_*.npz``.
What is NOT pinned by reference execution: DGL's conv arithmetic itself; it is
cross-checked against an independent dense formulation (``oracle.dense``),
hand-computed known-answer cases and fp64 gradcheck (tests/test_oracle_*.py).
"""
