"""Shared pieces of the application drivers (apps/*.py): plaintext slot packing, a device-resident
ciphertext with the reference's value semantics, and the set-up that every driver repeats.
Everything here goes through the C ABI (pyfhesi) and the C++ host layer's key generation; nothing
imports oracle/."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "fhe-si_b200"), os.path.join(ROOT, "scripts"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


# ---------------------------------------------------------------------------------------
# plaintext slots for p = 1 mod m (PlaintextSpace.cpp:22-134): slot j <-> root rho^(g^j)
# ---------------------------------------------------------------------------------------
class Slots:
    def __init__(self, m, p, g, phi):
        self.p, n = p, len(phi) - 1
        fs = [f for f in range(2, m + 1) if m % f == 0 and all(f % q for q in range(2, f))]
        x = 2
        while True:
            rho = pow(x, (p - 1) // m, p)
            if all(pow(rho, m // f, p) != 1 for f in fs):
                break
            x += 1
        self.roots, e = [], 1
        for _ in range(n):
            self.roots.append(pow(rho, e, p))
            e = e * g % m
        assert len(set(self.roots)) == n, "g does not generate Z_m^* (SURVEY.md §0.4)"
        self.total, self.usable = n, 1 << (n.bit_length() - 1)
        basis = np.zeros((n, n), dtype=np.int64)
        for j, r in enumerate(self.roots):
            b, carry = [0] * n, phi[n] % p
            for i in range(n - 1, -1, -1):
                b[i] = carry
                carry = (phi[i] + carry * r) % p
            d = 0
            for i in range(n - 1, -1, -1):
                d = (d * r + b[i]) % p
            di = pow(d, p - 2, p)
            basis[j] = [(v * di) % p for v in b]
        self.basis = basis

    def embed(self, values):
        """EmbedInSlots(msgs, onlyUsable=True): values -> polynomial coefficients mod p."""
        v = np.zeros(self.total, dtype=np.int64)
        v[:len(values)] = np.asarray(values, dtype=np.int64) % self.p
        return (v @ self.basis) % self.p

    def decode0(self, coeffs):
        acc = 0
        for c in reversed(list(coeffs)):
            acc = (acc * self.roots[0] + int(c)) % self.p
        return acc


# ---------------------------------------------------------------------------------------
# a device-resident ciphertext (value semantics, like the reference's Ciphertext)
# ---------------------------------------------------------------------------------------
class Ct:
    def __init__(self, env, buf, parts, scaled_up=False):
        self.env, self.buf, self.parts, self.scaled_up = env, buf, parts, scaled_up

    def copy(self):
        return Ct(self.env, self.buf.clone(), self.parts, self.scaled_up)

    def mul(self, other):  # Ciphertext::operator*=  -> tensor form
        e = self.env
        out = e.empty(e.dev.tprod_words(self.parts + other.parts - 1))
        e.dev.ct_tensor_dev(self.buf, self.parts, other.buf, other.parts, out, 1)
        return Ct(e, out, self.parts + other.parts - 1, True)

    def add_(self, other):
        assert self.scaled_up == other.scaled_up and self.parts == other.parts
        (self.env.dev.tprod_add_dev if self.scaled_up else self.env.dev.ct_add_dev)(self.buf, other.buf, self.parts, 1)
        return self

    def neg_(self):
        (self.env.dev.tprod_mul_scalar_dev if self.scaled_up else self.env.dev.ct_mul_scalar_dev)(self.buf, -1, self.parts, 1)
        return self

    def keyswitch_(self, ksw):  # KeySwitchSI::ApplyKeySwitch
        e = self.env
        if self.scaled_up:
            c = e.empty(e.dev.ct_words(self.parts))
            e.dev.scaledown_dev(self.buf, self.parts, c, 1)
            self.buf, self.scaled_up = c, False
        out = e.empty(e.dev.ct_words(2))
        e.dev.keyswitch_dev(ksw, self.buf, out, 1)
        self.buf, self.parts = out, 2
        return self

    def rotate_(self, k, ksw):  # tmp >>= k; autoKeySwitch.ApplyKeySwitch(tmp), one fused call
        e, d = self.env, self.env.dev
        assert self.parts == 2 and not self.scaled_up
        out = e.empty(d.ct_words(2))
        d.rotate_keyswitch_dev(ksw, self.buf, k, out, 1)
        self.buf = out
        return self


class Env:
    def __init__(self, dev, device, staging_bytes=0):
        self.dev, self.device = dev, device
        # pinned staging for the application's host -> device copies, made at start-up like the context itself
        # (a pageable cudaMemcpy goes through the driver's own bounce buffer and stalls for tens of
        # milliseconds every few calls on a busy host)
        self.pinned = None
        if staging_bytes and str(device).startswith("cuda"):
            self.pinned = torch.empty(int(staging_bytes), dtype=torch.uint8).pin_memory()

    def to_device(self, arr):
        """numpy array -> device tensor, through the pinned staging buffer when it fits."""
        arr = np.ascontiguousarray(arr)
        t = torch.from_numpy(arr)
        if self.pinned is None or arr.nbytes > self.pinned.numel():
            return t.to(self.device)
        stage = self.pinned[:arr.nbytes].view(t.dtype).view(t.shape)
        stage.copy_(t)
        out = torch.empty(t.shape, dtype=t.dtype, device=self.device)
        out.copy_(stage, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the staging buffer is reused by the next call
        return out

    def staging_array(self, shape, dtype=np.int32):
        """A zeroed numpy array that IS the pinned staging buffer (or an ordinary one when there is none / it is too
        small): what is packed into it goes up with upload_staged() without another host copy."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if self.pinned is None or nbytes > self.pinned.numel():
            return np.zeros(shape, dtype=dtype)
        a = self.pinned[:nbytes].numpy().view(dtype).reshape(shape)
        a.fill(0)
        return a

    def upload_staged(self, arr):
        """staging_array() -> device tensor."""
        if self.pinned is None or arr.nbytes > self.pinned.numel() or arr.ctypes.data != self.pinned.data_ptr():
            return self.to_device(arr)
        t = self.pinned[:arr.nbytes].view(torch.from_numpy(arr[:0]).dtype).view(arr.shape)
        out = torch.empty(t.shape, dtype=t.dtype, device=self.device)
        out.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the staging buffer is reused by the next call
        return out

    def empty(self, words):
        return torch.empty(int(words), dtype=torch.int32, device=self.device)




def sum_slots(ct, rot_k, rot_ksw):
    """SumBatchedData (Regression.h:166-178, Statistics.h:146-158): log2(usableSlots) rotate-and-add
    steps leave the sum of the usable slots in slot 0."""
    for kk, rk in zip(rot_k, rot_ksw):
        tmp = ct.copy().rotate_(kk, rk)
        ct.add_(tmp)
    return ct


def keyswitch_and_sum_slots_batch(env, tprods, count, ksw, rot_k, rot_ksw):
    """ApplyKeySwitch + SumBatchedData for `count` tensor-form accumulators at once
    (Regression.h:106-108,166-178): one scale-down, one key switch and log2(usableSlots)
    rotate-and-add steps over the whole batch instead of one call per ciphertext.
    tprods: [count][3 * Lt * N] words -> list of `count` 2-part Ct."""
    dev = env.dev
    cw = dev.ct_words(2)
    c3 = env.empty(count * dev.ct_words(3))
    dev.scaledown_dev(tprods, 3, c3, count)
    cur = env.empty(count * cw)
    dev.keyswitch_dev(ksw, c3, cur, count)
    tmp = env.empty(count * cw)
    for kk, rk in zip(rot_k, rot_ksw):
        dev.rotate_keyswitch_dev(rk, cur, kk, tmp, count)
        dev.ct_add_dev(cur, tmp, 2, count)
    cur = cur.view(count, cw)
    return [Ct(env, cur[i].clone(), 2) for i in range(count)]


def _tensor_sum_keyswitch(env, ksw, terms):
    """groups[g] = KeySwitch(ScaleDown(sum_t A_t[g] * B_t[g])): `terms` is a list of (A, B) pairs of
    [G][ct_words(2)] tensors; every product of a term is one batched tensor call, the sum is taken in tensor
    form (Matrix.cpp:80-97: products are summed before anything is reduced), then ONE ScaleDown and ONE key
    switch over the G groups.  Sums mod p_i are exact, so the result equals the reference's one-at-a-time order."""
    dev = env.dev
    G = terms[0][0].shape[0]
    tw = dev.tprod_words(3)
    acc = env.empty(G * tw)
    tmp = env.empty(G * tw) if len(terms) > 1 else None
    for t, (A, B) in enumerate(terms):
        dev.ct_tensor_dev(A.contiguous(), 2, B.contiguous(), 2, acc if t == 0 else tmp, G)
        if t:
            dev.tprod_add_dev(acc, tmp, 3, G)
    c3 = env.empty(G * dev.ct_words(3))
    dev.scaledown_dev(acc, 3, c3, G)
    out = env.empty(G * dev.ct_words(2)).view(G, -1)
    dev.keyswitch_dev(ksw, c3, out, G)
    return out


def _negated(env, cts):
    """-c for a batch in coefficient form (Ciphertext *= -1, Ciphertext.cpp:233-244)."""
    out = cts.clone()
    if out.shape[0]:
        env.dev.ct_mul_scalar_dev(out, -1, 2, out.shape[0])
    return out


def adjugate_and_det_batched(env, E, d, ksw):
    """Matrix<Ciphertext>::Invert (Matrix.cpp:181-213) with Determinant's Laplace expansion along the first free
    row (:223-262) and a key switch after every level -- the same circuit as the one-ciphertext-at-a-time
    recursion, evaluated level by level: all minors of one size in a handful of batched calls, each distinct
    minor once.  E: [d*d][ct_words(2)] entries, row-major.  -> (adj [d*d][cw] row-major, det [1][cw])."""
    idx = tuple(range(d))
    # minors needed, top-down: size s -> set of (rows, cols)
    need = {d - 1: {(tuple(r for r in idx if r != i), tuple(c for c in idx if c != j)) for i in idx for j in idx}}
    for s in range(d - 1, 1, -1):
        need[s - 1] = {(rows[1:], tuple(c for c in cols if c != col)) for rows, cols in need[s] for col in cols}
    val = {}  # (rows, cols) -> (tensor, row index in it)
    ent = lambda r, c: r * d + c
    for s in range(1, d):
        keys = sorted(need[s])
        if s == 1:
            buf = E[torch.tensor([ent(r[0], c[0]) for r, c in keys], device=E.device)]
        else:
            terms = []
            for t in range(s):
                left = E[torch.tensor([ent(rows[0], cols[t]) for rows, cols in keys], device=E.device)]
                if t % 2 == 1:
                    left = _negated(env, left)
                sub = [val[(rows[1:], tuple(c for c in cols if c != cols[t]))] for rows, cols in keys]
                src = sub[0][0]
                assert all(x[0] is src for x in sub)
                right = src[torch.tensor([x[1] for x in sub], device=E.device)]
                terms.append((left, right))
            buf = _tensor_sum_keyswitch(env, ksw, terms)
        for k, key in enumerate(keys):
            val[key] = (buf, k)
    # adj[j][i] = (-1)^(i+j) minor(without row i, col j)
    top, order = None, []
    for i in idx:
        for j in idx:
            b, k = val[(tuple(r for r in idx if r != i), tuple(c for c in idx if c != j))]
            top = b
            order.append((j * d + i, k, (i + j) % 2))
    order.sort()
    cof = top[torch.tensor([k for _, k, _ in order], device=E.device)]
    odd = torch.tensor([pos for pos, (_, _, sgn) in enumerate(order) if sgn], device=E.device)
    adj = cof.clone()
    if len(odd):
        adj[odd] = _negated(env, cof[odd])
    # det = sum_i M[0][i] * adj[i][0]   (Matrix.cpp:205-211)
    terms = [(E[ent(0, i):ent(0, i) + 1], adj[i * d:i * d + 1]) for i in idx]
    det = _tensor_sum_keyswitch(env, ksw, terms)
    return adj, det


def rotation_exponents(g, m, usable):
    """k = g, g^2, g^4, ... (Regression.h:70-81)."""
    out, k, ns = [], g % m, usable
    while ns > 1:
        out.append(k)
        ns >>= 1
        k = k * k % m
    return out


def embed_batch(env, slots, values):
    """values: int array [count][<= usable] -> uint32 message coefficients [count][n] on the device:
    PlaintextSpace::EmbedInSlots for a whole batch (fhesi_embed_slots_dev, exact integer arithmetic).
    The host only hands over the raw values; reduction mod p and the padding to the slot count happen
    on the device."""
    dev = env.dev
    cnt, width = values.shape
    if not hasattr(slots, "d_basis") or slots.d_basis.device != torch.device(env.device):
        slots.d_basis = torch.from_numpy(slots.basis.astype(np.int32)).to(env.device)
    d_msgs = torch.empty((max(cnt, 1), dev.n), dtype=torch.int32, device=env.device)
    if cnt:
        raw = env.upload_staged(np.ascontiguousarray(values, dtype=np.int32))
        d_vals = torch.zeros((cnt, slots.total), dtype=torch.int32, device=env.device)
        d_vals[:, :width] = torch.remainder(raw, slots.p).to(torch.int32)
        dev.embed_slots_dev(slots.d_basis, slots.total, d_vals, d_msgs, cnt)
    return d_msgs


def encryption_randomness(env, count, nrng):
    """r (uniform bits) and e (rounded Gaussians, sigma = 3.2) for `count` encryptions
    (FHE-SI.cpp:14-25).  On a GPU they are drawn on the device (torch's generator, seeded from
    `nrng`): no host sampling, no upload.  Host tensors (emulator tests) use numpy."""
    n = env.dev.n
    if str(env.device).startswith("cuda"):
        gen = torch.Generator(device=env.device)
        gen.manual_seed(int(nrng.integers(0, 2**62)))
        r_bits = torch.randint(0, 2, (count, n), dtype=torch.uint8, device=env.device, generator=gen)
        e = torch.round(torch.randn((count, 2, n), device=env.device, generator=gen) * 3.2).to(torch.int32)
        return r_bits, e
    r_bits = torch.from_numpy(nrng.integers(0, 2, size=(count, n), dtype=np.uint8)).to(env.device)
    e = torch.from_numpy(np.rint(nrng.normal(0.0, 3.2, size=(count, 2, n))).astype(np.int32)).to(env.device)
    return r_bits, e


def encrypt_batch(env, dpk, d_msgs, count, nrng):
    """FHESIPubKey::Encrypt over a batch with explicit randomness (FHE-SI.cpp:10-36)."""
    import os
    import sys
    import time
    dev = env.dev
    t0 = time.perf_counter()
    cts = env.empty(max(count, 1) * dev.ct_words(2)).view(max(count, 1), -1)
    if count:
        r_bits, e = encryption_randomness(env, count, nrng)
        t1 = time.perf_counter()
        dev.encrypt_dev(dpk, d_msgs, r_bits, e, cts, count)
        if os.environ.get("FHESI_APP_TIMING"):
            t2 = time.perf_counter()
            dev.sync()
            print(f"encrypt_batch({count}): randomness {t1 - t0:.4f}  enqueue {t2 - t1:.4f}  device {time.perf_counter() - t2:.4f}",
                  file=sys.stderr)
    return cts


def add_mask(env, dpk, slots, ct, nrng):
    """GenerateNoise (Regression.h:180-189): uniformly random values in every slot but the first."""
    v = np.zeros((1, slots.total), dtype=np.int64)
    v[0, 1:] = nrng.integers(0, slots.p, size=slots.total - 1)
    msg = ((v @ slots.basis) % slots.p).astype(np.int32)
    nz = encrypt_batch(env, dpk, torch.from_numpy(msg).to(env.device), 1, nrng)
    return ct.add_(Ct(env, nz[0].contiguous(), 2))


def decrypt_slot0(env, dsk, slots, ct):
    mbuf = torch.empty(env.dev.n, dtype=torch.int32, device=env.device)
    env.dev.decrypt_dev(dsk, ct.buf, 2, mbuf, 1)
    env.dev.sync()
    return slots.decode0(mbuf.cpu().numpy().view(np.uint32))
