"""Oracle for the VFM hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the shipped product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and only as the checker / the timed CPU
baseline -- never as a fallback for the CUDA path.

Contents
--------
``ref_slice``   AST-slices ``class CF`` out of the reference scripts under
                ``/root/reference`` at run time (this container only; the
                reference is not present on the GPU box) and drives it with
                the loss / backward / Adam lines of the scripts' loops.
``vfm_port``    fp32 torch restatement of the same algorithm (same ATen ops:
                ``torch.unique``, ``torch.distributions``, autograd, dense
                ``torch.optim.Adam``).  Travels to the GPU box; it is the
                checker in the ``-m gpu`` tests and the timed CPU baseline
                (``cpu_baseline.kind == "port"``).
``vfm_math``    independent fp64 numpy restatement with hand-derived
                gradients, used to cross-check the port
                and for F>2 pairwise sampling where the reference has no code.
``gen_golden``  writes ``tests/golden/*.npz`` from the sliced reference.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4, 8c), so the pin is the reference code itself executed in this
container: ``tests/test_oracle_vs_reference.py`` checks ``vfm_port`` and
``vfm_math`` against the sliced classes, and ``tests/golden/`` holds outputs
of the sliced classes on seeded inputs (generator committed).
"""
