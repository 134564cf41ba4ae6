"""CPU oracle for the CLUSTEN hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of the reference's algorithm
(Eiphodos/autofocusformerMod) used as the *checker*.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import it.  The product package ``autofocusformermod_b200`` never does, and has no CPU path.

Parity pinning status (see DESIGN.md "Oracle"):
  * QK / AV / WF / WEIGHTEDGATHER fwd+bwd: pinned against the reference's OWN CUDA kernels, built
    unmodified for sm_100a into ``oracle/_ref`` (``oracle/build_ref.py``) and run on the GPU box
    (``tests/test_gpu_vs_reference_kernels.py``), and against the reference's PyTorch gather
    formulation (``clusten/test_wg_kernel.py:38-39``, ``point_utils.py:116-117``).
  * space_filling_cluster / ClusterMerging selection / AFF backbone: pinned against outputs of
    the reference's own Python (``point_utils.py``, ``aff.py``) imported in the authoring
    container with stubs (``oracle/ref_loader.py``); vectors in ``tests/golden`` made by
    ``oracle/make_golden.py``.
  * kNN (pykeops 2.1.1, not vendored, not installable): PARITY UNPINNED for tie order and FMA
    contraction; canonical rule = IEEE fp32 without contraction, ties -> lowest database index.
  * torch.sort / torch.topk tie order (unstable / unspecified in the reference): PARITY UNPINNED;
    canonical rule = stable.
"""
