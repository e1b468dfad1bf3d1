"""CPU oracle for the cellranger-atac fragment -> peaks -> matrix path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: it may be
imported by ``tests/``, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` and by nothing
else.  The product path (``cellranger-atac_b200/``) never imports it and fails
loudly when the CUDA library is missing.

Parity pin: the reference's own tests hold no vector for this path (SURVEY.md
section 4).  The oracle is pinned instead against the reference's *own function
bodies executed in the build container* (``tests/golden/make_golden.py``
extracts them from ``/root/reference`` with ``ast`` at generation time and runs
them under Python 3 with ``xrange``/py2-comparison shims); the resulting
input/output vectors are committed under ``tests/golden/``.
"""
