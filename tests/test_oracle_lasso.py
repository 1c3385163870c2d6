"""CPU: the closed-form 2-atom non-negative LASSO (the restatement of spams.lasso, stain_utils.py:78) against an
independent coordinate-descent solver."""
import numpy as np
from sklearn.linear_model import Lasso

from oracle import stain_oracle as so


def test_closed_form_matches_sklearn():
    rng = np.random.default_rng(0)
    worst = 0.0
    for trial in range(300):
        D = rng.normal(size=(3, 2))
        if trial % 3 == 0:
            D = np.abs(D)
        D /= np.linalg.norm(D, axis=0)
        x = rng.normal(size=3) * (2.0 if trial % 2 else 0.2)
        if trial % 5 == 0:
            x = np.abs(x)
        lam = [0.01, 0.1, 0.5][trial % 3]
        a = so.lasso_pos2(x.reshape(3, 1), D, lam)[:, 0]
        # sklearn minimises (1/(2n))||x - Dw||^2 + alpha||w||_1 with n = 3 samples -> alpha = lam / 3
        sk = Lasso(alpha=lam / 3, positive=True, fit_intercept=False, tol=1e-14, max_iter=200000).fit(D, x).coef_
        worst = max(worst, np.abs(a - sk).max())
    assert worst < 1e-9, worst


def test_objective_is_minimal():
    rng = np.random.default_rng(1)
    D = np.abs(rng.normal(size=(3, 2)))
    D /= np.linalg.norm(D, axis=0)
    X = np.abs(rng.normal(size=(3, 200)))
    A = so.lasso_pos2(X, D, 0.1)

    def f(A_):
        R = X - D @ A_
        return 0.5 * (R * R).sum(0) + 0.1 * A_.sum(0)
    base = f(A)
    for _ in range(20):
        P = np.maximum(A + rng.normal(scale=1e-3, size=A.shape), 0)
        assert (f(P) >= base - 1e-12).all()
