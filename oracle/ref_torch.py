"""CPU oracle, second implementation (torch float64 + autograd).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see oracle/ref_numpy.py header): no reference test pins this path and the
reference's dependencies cannot be installed here.

This twin is written independently of oracle/ref_numpy.py and mirrors the *reference's own op
sequence and class structure* -- a Python loop over experts, one Cholesky + triangular solve per
GP (per sample under the reference's tf.map_fn), reverse-mode autodiff through everything -- so it
is also what bench.py times as the "restated reference" CPU baseline.
Gradients are taken w.r.t. the *unconstrained* variables exactly as the reference's
``tape.gradient(loss, trainable_variables)`` does.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import it.
"""
import math

import torch

DEFAULT_JITTER = 1e-6
DT = torch.float64


def _t(x):
    return torch.as_tensor(x, dtype=DT)


# -- bijectors (gpflow.utilities.positive / triangular) ------------------------------------------
def softplus(x):
    return torch.nn.functional.softplus(x, beta=1.0, threshold=1e9)


def softplus_inverse(y):
    y = _t(y)
    return y + torch.log(-torch.expm1(-y))


def fill_triangular_index(n):
    m = n * (n + 1) // 2
    ar = torch.arange(m)
    return torch.cat([ar[n:], ar.flip(0)]).reshape(n, n)


def fill_triangular(x):
    m = x.shape[-1]
    n = int(round(math.sqrt(0.25 + 2.0 * m) - 0.5))
    return torch.tril(x[..., fill_triangular_index(n)])


def fill_triangular_inverse(L):
    L = _t(L)
    n = L.shape[-1]
    idx = fill_triangular_index(n)
    ii, jj = torch.tril_indices(n, n)
    out = torch.zeros(L.shape[:-2] + (n * (n + 1) // 2,), dtype=DT)
    out[..., idx[ii, jj]] = L[..., ii, jj]
    return out


NOISE_LOWER = 1e-6  # mogpe/gps/likelihoods.py:21,40-42 (gpflow Gaussian default too)


class SVGP:
    """Stand-in for gpflow.models.SVGP(whiten=True, q_diag=False) with unconstrained leaves."""

    def __init__(self, gp, is_expert):
        Z = _t(gp["Z"])
        M, D = Z.shape
        self.Z = Z.clone().requires_grad_(True)
        self.ard = []
        self.ls_u, self.kvar_u = [], []
        self.active = []  # per kernel: the input columns it reads (gpflow active_dims) or None
        for k in gp["kernels"]:
            ad = k.get("active_dims")
            self.active.append(None if ad is None else torch.as_tensor(list(ad), dtype=torch.long))
            ls = _t(k["lengthscales"])
            self.ls_u.append(softplus_inverse(ls).clone().requires_grad_(True))
            self.kvar_u.append(softplus_inverse(_t(k["variance"])).clone().requires_grad_(True))
        self.q_mu = _t(gp["q_mu"]).clone().requires_grad_(True)
        self.L = self.q_mu.shape[1]
        self.q_diag = bool(gp.get("q_diag"))
        if self.q_diag:  # gpflow: q_sqrt [M, L] with the positive() (softplus) transform
            self.q_sqrt_u = softplus_inverse(_t(gp["q_sqrt"])).clone().requires_grad_(True)
        else:
            self.q_sqrt_u = fill_triangular_inverse(torch.tril(_t(gp["q_sqrt"]))).clone().requires_grad_(True)
        c = gp.get("mean_c")
        self.c = None if c is None else _t(c).reshape(-1).clone().requires_grad_(True)
        self.noise_u = None
        if is_expert:
            nv = _t(gp["noise_variance"]).reshape(-1)
            self.noise_u = softplus_inverse(nv - NOISE_LOWER).clone().requires_grad_(True)

    # parameter access in constrained space
    def kernel(self, l):
        i = 0 if len(self.ls_u) == 1 else l
        return softplus(self.ls_u[i]), softplus(self.kvar_u[i])

    @property
    def q_sqrt(self):
        if self.q_diag:
            return torch.diag_embed(softplus(self.q_sqrt_u).T)  # [L, M, M]
        return fill_triangular(self.q_sqrt_u)

    @property
    def noise_variance(self):
        return NOISE_LOWER + softplus(self.noise_u)

    def variables(self):
        """Named unconstrained leaves, in a fixed order."""
        out = [("Z", self.Z)]
        for i, (a, b) in enumerate(zip(self.ls_u, self.kvar_u)):
            out += [(f"kernels/{i}/lengthscales", a), (f"kernels/{i}/variance", b)]
        out += [("q_mu", self.q_mu), ("q_sqrt", self.q_sqrt_u)]
        if self.c is not None:
            out.append(("mean_c", self.c))
        if self.noise_u is not None:
            out.append(("noise_variance", self.noise_u))
        return out

    # gpflow arithmetic ------------------------------------------------------------------------
    def K(self, l, X, X2=None):
        ls, var = self.kernel(l)
        ad = self.active[0 if len(self.ls_u) == 1 else l]
        if ad is not None:  # gpflow Kernel.slice
            X = X[:, ad]
            X2 = None if X2 is None else X2[:, ad]
        Xs = X / ls
        if X2 is None:
            sq = (Xs * Xs).sum(-1, keepdim=True)
            dist = -2.0 * Xs @ Xs.T + (sq + sq.T)
        else:
            X2s = X2 / ls
            dist = -2.0 * Xs @ X2s.T + ((Xs * Xs).sum(-1)[:, None] + (X2s * X2s).sum(-1)[None, :])
        return var * torch.exp(-0.5 * dist)

    def mean_function(self, N):
        if self.c is None:
            return torch.zeros(N, self.L, dtype=DT)
        return self.c.reshape(1, -1).expand(N, self.L)

    def conditional(self, X, f, q_sqrt):
        """gpflow conditional(white=True, full_cov=False): f [M, L] -> mean [N, L], var [N, L]."""
        M = self.Z.shape[0]
        means, vars_ = [], []
        n_kern = len(self.ls_u)
        cache = {}
        for l in range(self.L):
            ki = 0 if n_kern == 1 else l
            if ki not in cache:
                Kmm = self.K(l, self.Z) + DEFAULT_JITTER * torch.eye(M, dtype=DT)
                Kmn = self.K(l, self.Z, X)
                Lm = torch.linalg.cholesky(Kmm)
                A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)
                cache[ki] = A
            A = cache[ki]
            fvar = self.kernel(l)[1] - (A * A).sum(0)
            means.append(A.T @ f[:, l])
            if q_sqrt is not None:
                LTA = torch.tril(q_sqrt[l]).T @ A
                fvar = fvar + (LTA * LTA).sum(0)
            vars_.append(fvar)
        return torch.stack(means, -1), torch.stack(vars_, -1)

    def predict_f(self, X):
        mu, var = self.conditional(X, self.q_mu, self.q_sqrt)
        return mu + self.mean_function(X.shape[0]), var

    def predict_f_samples(self, X, eps):
        """gpflow sample_mvn(full_cov=False): mean + sqrt(var) * eps, eps [S, N, L]."""
        mu, var = self.predict_f(X)
        return mu[None] + torch.sqrt(var)[None] * eps

    def predict_f_given_inducing_samples(self, X, eps_u):
        """mogpe/keras/gps.py:14-50: one conditional per sample (as tf.map_fn does)."""
        q_sqrt = self.q_sqrt
        u = self.q_mu.T[None] + torch.einsum("lmk,slk->slm", q_sqrt, eps_u)  # [S, L, M]
        means, vars_ = [], []
        for s in range(u.shape[0]):
            m, v = self.conditional(X, u[s].T, None)
            means.append(m)
            vars_.append(v)
        return torch.stack(means) + self.mean_function(X.shape[0])[None], torch.stack(vars_)

    def prior_kl(self):
        q_sqrt = self.q_sqrt
        M, L = self.q_mu.shape
        mahal = (self.q_mu**2).sum()
        logdet = torch.log(torch.diagonal(q_sqrt, dim1=-2, dim2=-1) ** 2).sum()
        trace = (q_sqrt**2).sum()
        return 0.5 * (mahal - M * L - logdet + trace)


def inv_probit(x):
    jitter = 1e-3
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) * (1 - 2 * jitter) + jitter


class MixtureOfSVGPExperts:
    """Both flavours of the reference model (mogpe/keras/mixture_of_experts/mosvgpe.py and
    mogpe/mixture_of_experts/svgp.py), restated over torch."""

    def __init__(self, spec):
        self.flavour = spec["flavour"]
        self.bound = spec["bound"]
        self.num_data = spec.get("num_data")
        self.experts = [SVGP(gp, True) for gp in spec["experts"]]
        self.gating = SVGP(spec["gating"], False)
        self.K = len(self.experts)
        self.G = self.gating.L

    def named_variables(self):
        out = []
        for k, e in enumerate(self.experts):
            out += [(f"experts/{k}/{n}", v) for n, v in e.variables()]
        out += [(f"gating/{n}", v) for n, v in self.gating.variables()]
        return out

    # -- gating ------------------------------------------------------------------------------
    def _probs_given_h(self, h):
        if self.G == 1:
            if self.flavour == "keras":
                return inv_probit(torch.cat([h, -h], -1))
            p = inv_probit(h)
            return torch.cat([p, 1 - p], -1)
        return torch.softmax(h, -1)

    def _probs_given_h_meanvar(self, hm, hv, eps_mc=None):
        if self.G == 1:
            if self.flavour == "keras":
                return inv_probit(torch.cat([hm, -hm], -1) / torch.sqrt(1 + torch.cat([hv, hv], -1)))
            p = inv_probit(hm / torch.sqrt(1 + hv))
            return torch.cat([p, 1 - p], -1)
        if eps_mc is None:
            raise ValueError("Softmax predict_mean_and_var is Monte-Carlo: pass eps_mc")
        return torch.softmax(hm[None] + torch.sqrt(hv)[None] * _t(eps_mc), -1).mean(0)

    def _h_samples(self, X, eps_h):
        hm, hv = self.gating.predict_f(X)
        sd = torch.sqrt(hv) if self.flavour == "keras" else hv
        return hm[None] + sd[None] * eps_h

    # -- experts -----------------------------------------------------------------------------
    def _expert_probs_inducing(self, X, Y, eps_u_experts):
        cols = []
        for k, e in enumerate(self.experts):
            fm, fv = e.predict_f_given_inducing_samples(X, _t(eps_u_experts[k]))
            yv = fv + e.noise_variance.expand(e.L)[None, None, :]
            sc = yv if self.flavour == "keras" else torch.sqrt(yv)
            cols.append(torch.distributions.Normal(fm, sc).log_prob(Y[None]).exp().prod(-1))
        return torch.stack(cols, -1)

    def _finish(self, experts_probs, mixing_probs, N):
        marg = torch.matmul(
            experts_probs[:, None, :, None, :], mixing_probs[None, :, :, :, None]
        )[..., 0, 0]
        var_exp = torch.log(marg).mean((0, 1)).sum(0)
        scale = 1.0 if self.num_data is None else self.num_data / N
        kl_g = self.gating.prior_kl()
        kl_e = sum(e.prior_kl() for e in self.experts)
        return var_exp * scale - kl_g - kl_e

    def elbo(self, X, Y, eps):
        X, Y = _t(X), _t(Y)
        N = X.shape[0]
        b = self.bound
        if b == "further_gating":
            mixing = self._probs_given_h(self._h_samples(X, _t(eps["h"])))
            return self._finish(self._expert_probs_inducing(X, Y, eps["u_experts"]), mixing, N)
        if b == "tight":
            hm, hv = self.gating.predict_f_given_inducing_samples(X, _t(eps["u_gating"]))
            mc = eps.get("mc")
            mixing = torch.stack(
                [self._probs_given_h_meanvar(hm[s], hv[s], None if mc is None else mc[s]) for s in range(hm.shape[0])]
            )
            return self._finish(self._expert_probs_inducing(X, Y, eps["u_experts"]), mixing, N)
        if b == "further":
            mixing = self._probs_given_h(self._h_samples(X, _t(eps["h"])))
            eps_f = _t(eps["f"])
            if self.flavour == "keras":
                cols = []
                for k, e in enumerate(self.experts):
                    f_s = e.predict_f_samples(X, eps_f[..., k])
                    noise = e.noise_variance.expand(e.L)[None, None, :]
                    cols.append(torch.distributions.Normal(f_s, noise).log_prob(Y[None]).exp().prod(-1))
                return self._finish(torch.stack(cols, -1), mixing, N)
            logps = []
            for k, e in enumerate(self.experts):
                fm, fv = e.predict_f(X)
                f_s = fm[None] + fv[None] * eps_f[..., k]
                logps.append(torch.distributions.Normal(f_s, e.noise_variance.expand(e.L)[None, None, :]).log_prob(Y[None]))
            w = torch.stack(logps, -1) + torch.log(mixing)[:, :, None, :]
            var_exp = torch.logsumexp(w, -1).sum(-1).mean(0).sum(0)
            scale = 1.0 if self.num_data is None else self.num_data / N
            return var_exp * scale - self.gating.prior_kl() - sum(e.prior_kl() for e in self.experts)
        raise ValueError(b)

    def elbo_and_grads(self, X, Y, eps):
        """-> (elbo float, {name: grad ndarray}) w.r.t. unconstrained variables."""
        named = self.named_variables()
        for _, v in named:
            v.grad = None
        val = self.elbo(X, Y, eps)
        grads = torch.autograd.grad(val, [v for _, v in named], allow_unused=True)
        out = {}
        for (n, v), g in zip(named, grads):
            out[n] = (torch.zeros_like(v) if g is None else g).detach().numpy()
        return float(val.detach()), out

    # -- prediction --------------------------------------------------------------------------
    def predict_mixing_probs(self, X, eps_mc=None):
        hm, hv = self.gating.predict_f(_t(X))
        return self._probs_given_h_meanvar(hm, hv, eps_mc)

    def predict_experts_y(self, X):
        X = _t(X)
        mus, vs = [], []
        for e in self.experts:
            fm, fv = e.predict_f(X)
            mus.append(fm)
            vs.append(fv + e.noise_variance.expand(e.L)[None, :])
        return torch.stack(mus, -1), torch.stack(vs, -1)

    def predict_y(self, X, eps_mc=None):
        pi = self.predict_mixing_probs(X, eps_mc)[:, None, :]
        mu, var = self.predict_experts_y(X)
        mean = (pi * mu).sum(-1)
        if self.flavour == "keras":
            mv = (pi * var * var).sum(-1) + (pi * mu * mu).sum(-1) - mean * mean
            mv = torch.sqrt(mv) ** 2
        else:
            mv = (pi * var).sum(-1) + (pi * (mu - mean[..., None]) ** 2).sum(-1)
        return mean, mv
