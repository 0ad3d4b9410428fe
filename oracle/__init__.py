"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the differentiable-WDF hot path. Importable only from tests/, from
``__graft_entry__.smoke()`` and from ``bench.py``'s cpu_baseline / ``--impl reference`` legs — as
the checker or the timed CPU baseline, never as the product path. The product package
(``differentiable-wdfs_b200``) neither imports nor links anything in this directory.

* ``oracle.cpu``       ctypes bindings over ``oracle/_ref/libdwdf_oracle.so`` (the plain-C
                       restatement, ``wdf_oracle.c``) and ``oracle/_ref/libdwdf_ref.so`` (the
                       unmodified reference C++ compiled in place, ``ref_harness.cpp``).
* ``oracle.torch_wdf`` line-for-line torch restatement of ``wdf_py/lib/tf_wdf.py`` (TensorFlow 2.5
                       is not installable here) + an analytic DiodePair root; gradient oracle via
                       ``torch.autograd`` and the stand-in for the wdf_py TensorFlow CPU path.
* ``oracle.nn``        numpy / torch restatement of the neural-root clipper (``layers.py``, ``clipper_pot.py:94-127``).
* ``oracle/shim_tf``   a minimal ``tensorflow`` look-alike on torch, so that ``tests/golden/make_golden_py_reference.py``
                       can execute the UNMODIFIED reference Python sources and record their outputs and
                       ``tape.gradient`` results (``tests/golden/py_reference_vectors.npz``): what pins the two
                       restatements above, and the kernels, to the reference's own code — gradients included.
"""
