"""CPU restatement of the OUTER network step around the ODE block — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's MNIST model (examples/mnist/OdeNetModel.java:60-88) is a DL4J ComputationGraph:

    input [B,1,28,28]
      -> ResBlockStem (examples/mnist/ResBlockStem.java:25-48): inputConv 3x3 (Truncate, bias) -> 2 x { BN-ReLU ; [1x1 stride 2]  +
                                                               [3x3 stride 2 Same ReLU -> BN-ReLU -> 3x3 Same] }          28 -> 26 -> 13 -> 7
      -> OdeVertex (the hot path; oracle/node_oracle.cpp)
      -> Output (examples/mnist/Output.java:27-35): BN-ReLU -> global average pooling -> Dense(10) softmax + LossMCXENT
    updater: Nesterovs(momentum 0.9), learning rate by epoch (examples/mnist/LayerUtil.java:30-47); BatchNorm running mean / var are
    "parameters without gradient" updated through the gradient view (ode/vertex/impl/helper/OdeGraphHelper.java:105-119).

All of that arithmetic lives in DL4J / ND4J 1.0.0-beta3 (third party, not vendored, cannot run here): this file restates the layers'
published definitions with numpy — fp32 tensors, every reduction accumulated in float64 and rounded once (the arithmetic contract of
DESIGN.md section 2) — and is tied to an independent implementation (torch autograd, fp64) in tests/test_outer_net.py.
PARITY UNPINNED by the reference itself (it has no test with numbers for these layers).  Conventions restated:

  * ConvolutionLayer: cross-correlation, weights [Cout,Cin,kh,kw] c-order, flat parameters = bias (if any) FIRST, then weights
    (ConvolutionParamInitializer); Truncate: out = (in - k)/s + 1; Same: out = ceil(in/s), total padding (out-1)*s + k - in, the
    smaller half in front (top / left).
  * BatchNormalization (training): batch mean, biased batch variance, y = gamma*(x-mean)/sqrt(var+eps)+beta, eps 1e-5, decay 0.9;
    flat parameters gamma, beta, mean, var; running statistics move as mean <- decay*mean + (1-decay)*batch_mean (same for var).
  * GlobalPoolingLayer(AVG): mean over H, W.   OutputLayer: Dense W [nIn,nOut] f-order then bias; softmax with max subtraction;
    LossMCXENT score = -sum(labels*log(clip(p, 1e-10, 1-1e-10)))/B; gradient w.r.t. the pre-activations = p - labels.
  * gradients are summed over the minibatch and divided by B in the updater; Nesterovs: v' = mu*v - lr*g; p += -mu*v + (1+mu)*v'.
  * parameter order = topological order of the graph, ties in insertion order: inputConv, then per residual block firstNorm, downSample,
    firstConv, secondNorm, secondConv; the ODE block (its own flat view); normOutput; output.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
BN_EPS, BN_DECAY = 1e-5, 0.9


def _r(x):
    return np.asarray(x, dtype=np.float64).astype(F32)


def same_pad(n_in, k, s):
    n_out = -(-n_in // s)
    total = max((n_out - 1) * s + k - n_in, 0)
    return n_out, total // 2


def conv_fwd(x, w, b, stride, same, relu):
    """x [B,Ci,H,W], w [Co,Ci,kh,kw] -> (y, pre-activation kept implicit: y > 0 is the ReLU mask)."""
    B, Ci, H, W = x.shape
    Co, _, kh, kw = w.shape
    if same:
        Ho, pt = same_pad(H, kh, stride)
        Wo, pl = same_pad(W, kw, stride)
    else:
        Ho, pt, Wo, pl = (H - kh) // stride + 1, 0, (W - kw) // stride + 1, 0
    xp = np.zeros((B, Ci, H + kh + stride, W + kw + stride), dtype=np.float64)
    xp[:, :, pt:pt + H, pl:pl + W] = x
    acc = np.zeros((B, Co, Ho, Wo), dtype=np.float64)
    w64 = w.astype(np.float64)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, :, i:i + stride * Ho:stride, j:j + stride * Wo:stride]
            acc += np.einsum("bchw,oc->bohw", patch, w64[:, :, i, j], optimize=True)
    y = acc.astype(F32)
    if b is not None:
        y = (y + b.reshape(1, -1, 1, 1).astype(F32)).astype(F32)
    if relu:
        y = np.where(y > 0, y, F32(0)).astype(F32)
    return y, (pt, pl)


def conv_bwd(x, w, y, dy, stride, pad, relu, has_bias):
    """-> dx, dw, db (sums over the minibatch)."""
    B, Ci, H, W = x.shape
    Co, _, kh, kw = w.shape
    pt, pl = pad
    Ho, Wo = y.shape[2], y.shape[3]
    dz = np.where(y > 0, dy, F32(0)).astype(F32) if relu else dy.astype(F32)
    dz64 = dz.astype(np.float64)
    xp = np.zeros((B, Ci, H + kh + stride, W + kw + stride), dtype=np.float64)
    xp[:, :, pt:pt + H, pl:pl + W] = x
    dxp = np.zeros_like(xp)
    dw = np.zeros(w.shape, dtype=np.float64)
    w64 = w.astype(np.float64)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, :, i:i + stride * Ho:stride, j:j + stride * Wo:stride]
            dw[:, :, i, j] = np.einsum("bohw,bchw->oc", dz64, patch, optimize=True)
            dxp[:, :, i:i + stride * Ho:stride, j:j + stride * Wo:stride] += np.einsum("bohw,oc->bchw", dz64, w64[:, :, i, j], optimize=True)
    dx = dxp[:, :, pt:pt + H, pl:pl + W].astype(F32)
    db = dz64.sum(axis=(0, 2, 3)).astype(F32) if has_bias else None
    return dx, dw.astype(F32), db


def bn_fwd(x, gamma, beta, relu=True):
    m = x.astype(np.float64).mean(axis=(0, 2, 3))
    v = np.maximum((x.astype(np.float64) ** 2).mean(axis=(0, 2, 3)) - m * m, 0.0)
    mean, invstd = m.astype(F32), (1.0 / np.sqrt(v + BN_EPS)).astype(F32)
    xh = ((x - mean.reshape(1, -1, 1, 1)).astype(F32) * invstd.reshape(1, -1, 1, 1)).astype(F32)
    y = ((gamma.reshape(1, -1, 1, 1) * xh).astype(F32) + beta.reshape(1, -1, 1, 1)).astype(F32)
    if relu:
        y = np.where(y > 0, y, F32(0)).astype(F32)
    return y, (xh, invstd, mean, v.astype(F32))


def bn_bwd(y, dy, gamma, cache, relu=True):
    xh, invstd, _, _ = cache
    n = y.shape[0] * y.shape[2] * y.shape[3]
    dz = np.where(y > 0, dy, F32(0)).astype(F32) if relu else dy.astype(F32)
    s1 = dz.astype(np.float64).sum(axis=(0, 2, 3))
    s2 = (dz.astype(np.float64) * xh.astype(np.float64)).sum(axis=(0, 2, 3))
    dgamma, dbeta = s2.astype(F32), s1.astype(F32)
    m1, m2 = (s1 / n).astype(F32).reshape(1, -1, 1, 1), (s2 / n).astype(F32).reshape(1, -1, 1, 1)
    k = (gamma * invstd).astype(F32).reshape(1, -1, 1, 1)
    dx = (k * ((dz - m1).astype(F32) - (xh * m2).astype(F32)).astype(F32)).astype(F32)
    return dx, dgamma, dbeta


class OuterLayout:
    """Flat parameter layout of the whole model (DL4J order, see the module header)."""

    def __init__(self, C, n_ode):
        self.C = C
        off = 0
        self.slices = {}

        def add(name, n):
            nonlocal off
            self.slices[name] = (off, n)
            off += n
        add("inputConv.b", C); add("inputConv.W", C * 1 * 9)
        for r in (0, 1):
            add(f"firstNorm{r}", 4 * C)
            add(f"downSample{r}.W", C * C)
            add(f"firstConv{r}.W", C * C * 9)
            add(f"secondNorm{r}", 4 * C)
            add(f"secondConv{r}.W", C * C * 9)
        add("ode", n_ode)
        add("normOutput", 4 * C)
        add("output.W", C * 10); add("output.b", 10)
        self.n = off

    def get(self, p, name):
        o, n = self.slices[name]
        return p[o:o + n]

    def bn_names(self):
        return ["firstNorm0", "secondNorm0", "firstNorm1", "secondNorm1", "normOutput"]

    def noop_mask(self):
        """True where the updater is NoOp (BatchNorm running mean / var of the OUTER network)."""
        m = np.zeros(self.n, dtype=bool)
        for nm in self.bn_names():
            o, _ = self.slices[nm]
            m[o + 2 * self.C:o + 4 * self.C] = True
        return m


class OuterOracle:
    """One training step (forward, backward, update) of the MNIST ODE-net; the ODE block runs on oracle.Oracle."""

    def __init__(self, ode_oracle, ode_time, C, batch):
        self.ode, self.t, self.C, self.B = ode_oracle, np.asarray(ode_time, dtype=F32), C, batch
        self.layout = OuterLayout(C, int(ode_oracle.n_params))

    # -- forward through the stem, the ODE block and the output block; keeps what the backward pass needs --------------------
    def forward(self, p, images):
        L, C = self.layout, self.C
        g = lambda nm: L.get(p, nm)
        cache = {}
        W_in = g("inputConv.W").reshape(C, 1, 3, 3)
        a, _ = conv_fwd(images.astype(F32), W_in, g("inputConv.b"), 1, False, False)
        cache["img"] = images.astype(F32)
        for r in (0, 1):
            bn1 = g(f"firstNorm{r}")
            n1, c1 = bn_fwd(a, bn1[:C], bn1[C:2 * C])
            Wd = g(f"downSample{r}.W").reshape(C, C, 1, 1)
            ds, pd = conv_fwd(n1, Wd, None, 2, False, False)
            W1 = g(f"firstConv{r}.W").reshape(C, C, 3, 3)
            f1, p1 = conv_fwd(n1, W1, None, 2, True, True)
            bn2 = g(f"secondNorm{r}")
            n2, c2 = bn_fwd(f1, bn2[:C], bn2[C:2 * C])
            W2 = g(f"secondConv{r}.W").reshape(C, C, 3, 3)
            f2, p2 = conv_fwd(n2, W2, None, 1, True, False)
            cache[r] = dict(a_in=a, n1=n1, c1=c1, pd=pd, f1=f1, p1=p1, n2=n2, c2=c2, f2=f2, p2=p2)
            a = (ds + f2).astype(F32)
        cache["z0"] = a
        z1 = self.ode.forward(g("ode"), a, self.t)
        cache["ode_fwd_stats"] = (int(self.ode.stats.nfe), int(self.ode.stats.accepted), int(self.ode.stats.rejected))
        cache["z1"] = z1
        bno = g("normOutput")
        no, co = bn_fwd(z1, bno[:C], bno[C:2 * C])
        pool = no.astype(np.float64).mean(axis=(2, 3)).astype(F32)
        Wo = g("output.W").reshape(C, 10, order="F")
        logits = ((pool.astype(np.float64) @ Wo.astype(np.float64)).astype(F32) + g("output.b")).astype(F32)
        e = np.exp((logits - logits.max(axis=1, keepdims=True)).astype(F32)).astype(F32)
        probs = (e / e.astype(np.float64).sum(axis=1, keepdims=True).astype(F32)).astype(F32)
        cache.update(no=no, co=co, pool=pool, probs=probs)
        return probs, cache

    def score(self, probs, labels):
        pc = np.clip(probs.astype(np.float64), 1e-10, 1 - 1e-10)
        return float(-(labels.astype(np.float64) * np.log(pc)).sum() / self.B)

    def backward(self, p, labels, cache):
        """-> flat gradient (sums over the minibatch; BatchNorm mean/var slots hold the running-statistics 'gradient')."""
        L, C = self.layout, self.C
        g = lambda nm: L.get(p, nm)
        grad = np.zeros(L.n, dtype=F32)
        put = lambda nm, v: grad.__setitem__(slice(L.slices[nm][0], L.slices[nm][0] + L.slices[nm][1]), np.asarray(v, dtype=F32).ravel())

        def put_bn(nm, dgamma, dbeta, bn_cache):
            o, _ = L.slices[nm]
            grad[o:o + C], grad[o + C:o + 2 * C] = dgamma, dbeta
            _, _, bmean, bvar = bn_cache
            run = L.get(p, nm)
            grad[o + 2 * C:o + 3 * C] = ((run[2 * C:3 * C] - bmean).astype(F32) * F32(1 - BN_DECAY)).astype(F32)
            grad[o + 3 * C:o + 4 * C] = ((run[3 * C:4 * C] - bvar).astype(F32) * F32(1 - BN_DECAY)).astype(F32)

        probs, pool, no = cache["probs"], cache["pool"], cache["no"]
        dlog = (probs - labels.astype(F32)).astype(F32)
        Wo = g("output.W").reshape(C, 10, order="F")
        put("output.W", (pool.astype(np.float64).T @ dlog.astype(np.float64)).astype(F32).ravel(order="F"))
        put("output.b", dlog.astype(np.float64).sum(axis=0))
        dpool = (dlog.astype(np.float64) @ Wo.astype(np.float64).T).astype(F32)
        hw = no.shape[2] * no.shape[3]
        dno = np.broadcast_to((dpool / F32(hw)).astype(F32)[:, :, None, None], no.shape).astype(F32)
        bno = g("normOutput")
        dz1, dga, dbe = bn_bwd(no, dno, bno[:C], cache["co"])
        put_bn("normOutput", dga, dbe, cache["co"])
        o_ode, n_ode = L.slices["ode"]
        gview = np.zeros(n_ode, dtype=F32)            # ZeroGrad listener: the theta-adjoint is seeded from a zeroed view
        dz0, gode, _ = self.ode.backward(g("ode"), cache["z1"], dz1, self.t, grad_inout=gview)
        cache["ode_bwd_stats"] = (int(self.ode.stats.nfe), int(self.ode.stats.accepted), int(self.ode.stats.rejected))
        grad[o_ode:o_ode + n_ode] = gode
        da = dz0
        for r in (1, 0):
            c = cache[r]
            W2 = g(f"secondConv{r}.W").reshape(C, C, 3, 3)
            dn2, dW2, _ = conv_bwd(c["n2"], W2, c["f2"], da, 1, c["p2"], False, False)
            put(f"secondConv{r}.W", dW2)
            bn2 = g(f"secondNorm{r}")
            df1, dga, dbe = bn_bwd(c["n2"], dn2, bn2[:C], c["c2"])
            put_bn(f"secondNorm{r}", dga, dbe, c["c2"])
            W1 = g(f"firstConv{r}.W").reshape(C, C, 3, 3)
            dn1a, dW1, _ = conv_bwd(c["n1"], W1, c["f1"], df1, 2, c["p1"], True, False)
            put(f"firstConv{r}.W", dW1)
            Wd = g(f"downSample{r}.W").reshape(C, C, 1, 1)
            ds_out = np.zeros((self.B, C, da.shape[2], da.shape[3]), dtype=F32)   # identity activation: y is not needed
            dn1b, dWd, _ = conv_bwd(c["n1"], Wd, ds_out, da, 2, c["pd"], False, False)
            put(f"downSample{r}.W", dWd)
            dn1 = (dn1b + dn1a).astype(F32)           # epsilons of the two consumers of firstNorm are summed (downSample first, then firstConv)
            bn1 = g(f"firstNorm{r}")
            da, dga, dbe = bn_bwd(c["n1"], dn1, bn1[:C], c["c1"])
            put_bn(f"firstNorm{r}", dga, dbe, c["c1"])
        W_in = g("inputConv.W").reshape(C, 1, 3, 3)
        y_in = cache[0]["a_in"]
        _, dWin, dbin = conv_bwd(cache["img"], W_in, y_in, da, 1, (0, 0), False, True)
        put("inputConv.W", dWin)
        put("inputConv.b", dbin)
        return grad

    def update(self, p, v, grad, lr, mu):
        """Nesterovs on everything except the BatchNorm running statistics (NoOp updater: the 'gradient' is subtracted as it is)."""
        noop = self.layout.noop_mask()
        # the ODE block's own BatchNorm mean / var slots: gradient view untouched by the native path (zero) -> unchanged
        gB = (grad / F32(self.B)).astype(F32)
        v_new = ((F32(mu) * v).astype(F32) - (F32(lr) * gB).astype(F32)).astype(F32)
        step = ((F32(-mu) * v).astype(F32) + (F32(1 + mu) * v_new).astype(F32)).astype(F32)
        p_new = np.where(noop, (p - grad).astype(F32), (p + step).astype(F32)).astype(F32)
        v_out = np.where(noop, v, v_new).astype(F32)
        return p_new, v_out

    def train_step(self, p, v, images, labels, lr, mu=0.9):
        probs, cache = self.forward(p, images)
        sc = self.score(probs, labels)
        grad = self.backward(p, labels, cache)
        p2, v2 = self.update(p, v, grad, lr, mu)
        return dict(score=sc, probs=probs, grad=grad, params=p2, updater_state=v2, ode_fwd_stats=cache["ode_fwd_stats"],
                    ode_bwd_stats=cache["ode_bwd_stats"], z0=cache["z0"], z1=cache["z1"])
