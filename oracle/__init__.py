"""CPU oracle for the Flowtron hot path (TEST INFRASTRUCTURE ONLY).

This package is a CPU restatement (torch fp32 / numpy) of the reference
algorithm for the AR-flow hot path, the NLL loss and the mel front-end.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
package ``flowtron_b200`` never imports anything from here.

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the *reference itself* run in the build
container (``tests/make_golden.py`` imports ``/root/reference`` with the two
CPU shims of SURVEY.md §8c and writes ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks this restatement against those fixtures.
The mel filterbank (librosa, not in the tree) is restated from its published
algorithm; that boundary is "parity unpinned" by the reference's own tests.
"""
