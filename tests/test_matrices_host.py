"""Host-side semantics of the GeneralMatrix mirror (values, str(), exception classes and messages) against outputs of the
reference's algebra/matrices.py recorded by tests/golden/make_matrices_golden.py.  Plain-int elements: no GPU involved."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def run_cases(G):
    out = []
    m = G([[1, 2], [3, 4]])
    ops = [lambda: str(m), lambda: str(m + G([[1, 1], [1, 1]])), lambda: str(m * G([[1, 0], [0, 1]])), lambda: m == 0, lambda: G([[0, 0]]) == 0,
           lambda: str(-m), lambda: str(m % 3), lambda: str(m * 5), lambda: str(m - G([[1, 1], [1, 1]])), lambda: m == G([[1, 2], [3, 4]]),
           lambda: m == G([[1, 2]]), lambda: str(0 + m), lambda: len(m), lambda: m[1], lambda: m + G([[1, 2]]), lambda: m * G([[1, 2]]), lambda: m * 2.5,
           lambda: m + 2.5, lambda: m % 1, lambda: m % 2.0, lambda: m.norm("infty"), lambda: m.weight(),
           lambda: str(G([[1, 2, 3]]) * G([[1], [2], [3]])), lambda: str(sum([m, m, m]))]
    for f in ops:
        try:
            out.append(repr(f()))
        except Exception as e:  # the class and the message are part of the interface
            out.append(type(e).__name__ + ": " + str(e))
    for bad in ((), [], [1], [[]], [[1], [1, 2]], [[object]], [[1, 2.0]]):
        try:
            G(bad)
            out.append("ok")
        except Exception as e:
            out.append(type(e).__name__ + ": " + str(e))
    return out


def test_general_matrix_host_semantics_match_reference():
    from fusion_cryptography_b200.algebra.matrices import GeneralMatrix

    want = json.load(open(os.path.join(HERE, "golden", "matrices_host_golden.json")))
    assert run_cases(GeneralMatrix) == want
