"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/field.h).

Second, independent restatement of the reference's PCS hot path in pure Python
(big ints + hashlib).  It exists to cross-check the C oracle (oracle/oracle.c)
and to generate the golden fixtures under tests/golden/ (tests/golden/gen_golden.py).
Pure-Python loops: small sizes only.  Citations are relative to /root/reference/.
"""
import hashlib

M = 340282366920938463463374557953744961537  # src/ntt/mod.rs:35
LOG_BLOWUP = 1      # src/fri/mod.rs:16
NUM_QUERIES = 128   # src/fri/mod.rs:17


def new(x):  # BaseElement::new: one conditional subtraction
    return x - M if x >= M else x


def from_i64(v):  # src/field.rs:144-154: `val as u128` sign-extends
    return new(v & ((1 << 128) - 1))


def inv(a):
    return 0 if a == 0 else pow(a, M - 2, M)


def div(a, b):
    return a * inv(b) % M


def to_bytes(x):  # src/field.rs:33-38
    return x.to_bytes(16, "little")


def pow_2_generator(log_size):  # src/ntt/mod.rs:42-54
    m1 = M - 1
    max_log = (m1 & -m1).bit_length() - 1
    if log_size > max_log:
        return None
    return pow(3, m1 >> log_size, M)


def pow_2_generator_powers(log_size):  # src/ntt/mod.rs:18-28
    g = pow_2_generator(log_size)
    if g is None:
        return None
    out, cur = [], 1
    for _ in range(1 << log_size):
        out.append(cur)
        cur = cur * g % M
    return out


def bit_reverse_permutation(v):  # src/ntt/mod.rs:113-123
    n = len(v)
    bits = (n & -n).bit_length() - 1 if n else 64
    for i in range(n):
        j = int(format(i, "064b")[::-1], 2) >> (64 - bits) if bits else i
        if i < j:
            v[i], v[j] = v[j], v[i]


def _network(values, gen):  # src/ntt/mod.rs:76-109
    n = len(values)
    bit_reverse_permutation(values)
    for i in range(0, n, 2):
        u, v = values[i], values[i + 1]
        values[i], values[i + 1] = (u + v) % M, (u - v) % M
    ln = 4
    while ln <= n:
        cg = pow(gen, n // ln, M)
        gp, acc = [], 1
        for _ in range(ln // 2):
            gp.append(acc)
            acc = acc * cg % M
        for i in range(0, n, ln):
            for j in range(ln // 2):
                v = values[i + j + ln // 2] * gp[j] % M
                u = values[i + j]
                values[i + j] = (u + v) % M
                values[i + j + ln // 2] = (u - v) % M
        ln *= 2
    return values


def ntt(coeffs, gen):
    assert len(coeffs) & (len(coeffs) - 1) == 0 and coeffs
    return _network(list(coeffs), gen)


def intt(evals, gen):  # src/ntt/mod.rs:132-173
    n = len(evals)
    out = _network(list(evals), div(1, gen))
    n_inv = div(1, from_i64(n))
    return [x * n_inv % M for x in out]


def reed_solomon(coeffs, gen):  # src/fri/mod.rs:19-28
    return ntt(list(coeffs) + [0] * len(coeffs), gen)


def to_coefficient(evals):  # src/polynomials.rs:150-163
    c = list(evals)
    n = (len(c) & -len(c)).bit_length() - 1 if c else 0
    for i in range(n):
        mask = 1 << i
        for j in range(1 << n):
            if j & mask:
                c[j] = (c[j] - c[j ^ mask]) % M
    return c


def to_evaluation(coeffs):  # src/polynomials.rs:111-124
    c = list(coeffs)
    n = (len(c) & -len(c)).bit_length() - 1 if c else 0
    for i in range(n):
        mask = 1 << i
        for j in range(1 << n):
            if j & mask:
                c[j] = (c[j] + c[j ^ mask]) % M
    return c


def mle_evals_evaluate(evals, args):  # src/polynomials.rs:165-187
    acc = 0
    for pos, e in enumerate(evals):
        term = e
        for b, arg in enumerate(reversed(args)):
            term = term * (arg if (pos >> b) & 1 else (1 - arg) % M) % M
        acc = (acc + term) % M
    return acc


def mle_coeffs_evaluate(coeffs, args):  # src/polynomials.rs:126-146
    acc = 0
    for pos, c in enumerate(coeffs):
        term = c
        for b, arg in enumerate(reversed(args)):
            if (pos >> b) & 1:
                term = term * arg % M
        acc = (acc + term) % M
    return acc


def poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % M
    return acc


def interpolate(evals):  # src/polynomials.rs:51-86
    n = len(evals)
    coeffs = [0] * n
    for j, yj in enumerate(evals):
        lj, denom = [1], 1
        for m in range(n):
            if m == j:
                continue
            nl = [0] * (len(lj) + 1)
            for i, a in enumerate(lj):
                nl[i] = (nl[i] - a * m) % M
                nl[i + 1] = (nl[i + 1] + a) % M
            lj = nl
            denom = denom * (j - m) % M
        scale = div(yj, denom)
        for i in range(n):
            coeffs[i] = (coeffs[i] + scale * lj[i]) % M
    return coeffs


class Transcript:  # src/transcript.rs
    def __init__(self):
        self.state = hashlib.sha256()

    def clone(self):
        t = Transcript()
        t.state = self.state.copy()
        return t

    def absorb(self, b):
        self.state.update(bytes(b))

    def random(self):
        return self.state.copy().digest()

    def next_challenge(self):
        return new(int.from_bytes(self.random()[:16], "little"))


def hash_node(l, r):
    return hashlib.sha256(l + r).digest()


class Merkle:  # src/merkle_tree/mod.rs
    def __init__(self, layers, data, batched=False):
        self.layers, self.data, self.batched = layers, data, batched

    @staticmethod
    def _build(first):
        layers = [first]
        while len(layers[-1]) > 1:
            cur = layers[-1]
            layers.append([hash_node(cur[i], cur[i + 1]) for i in range(0, len(cur), 2)])
        return layers

    @staticmethod
    def commit(data):  # :65-85, items are bytes
        assert data and len(data) & (len(data) - 1) == 0
        return Merkle(Merkle._build([hashlib.sha256(d).digest() for d in data]), data)

    @staticmethod
    def batch_commit(data):  # :92-131
        n = len(data[0])
        assert n and n & (n - 1) == 0 and all(len(d) == n for d in data)
        first = [hashlib.sha256(b"".join(d[i] for d in data)).digest() for i in range(n)]
        return Merkle(Merkle._build(first), data, True)

    def root(self):
        return self.layers[-1][0]

    def open(self, index):  # :31-58 / :134-175 -> (value, [(digest, dir)])
        n = len(self.data[0]) if self.batched else len(self.data)
        if index >= n:
            return None
        value = b"".join(d[index] for d in self.data) if self.batched else self.data[index]
        path, cur = [], index
        for layer in self.layers:
            sib, d = (cur + 1, 1) if cur % 2 == 0 else (cur - 1, 0)
            if sib >= len(layer):
                break
            path.append((layer[sib], d))
            cur //= 2
        return value, path


def path_verify(value, path, root, index):  # :216-246
    h, computed = hashlib.sha256(value).digest(), 0
    for i, (sib, d) in enumerate(path):
        if d == 0:
            computed += 1 << i
            h = hash_node(sib, h)
        else:
            h = hash_node(h, sib)
    return h == root and computed == index


def commit_rs_code(code):  # src/fri/mod.rs:45-55
    h = len(code) // 2
    return Merkle.commit([to_bytes(code[i]) + to_bytes(code[i + h]) for i in range(h)])


class FriProverData:  # src/fri/mod.rs:57-175
    def __init__(self):
        self.trees, self.codes, self.last_element = [], [], None

    @staticmethod
    def init(code, t):
        f = FriProverData()
        m = commit_rs_code(code)
        f.trees.append(m)
        f.codes.append(list(code))
        t.absorb(m.root())
        return f

    def finish(self, nxt, t):
        if len(nxt) == 2:
            assert nxt[0] == nxt[1], "not an RS code"
            self.last_element = nxt[0]
            t.absorb(to_bytes(nxt[0]))
            return
        m = commit_rs_code(nxt)
        self.trees.append(m)
        self.codes.append(nxt)
        t.absorb(m.root())

    def fold_step(self, gen_pows, k, r, t):  # :79-134
        cur = self.codes[-1]
        n = len(cur)
        if n <= 2:
            return
        h = n // 2
        half = div(1, 2)
        nxt = []
        for i in range(h):
            a, b = cur[i], cur[i + h]
            w = 1 if i == 0 else gen_pows[len(gen_pows) - i * (1 << k)]
            nxt.append(((a + b) + r * ((a - b) * w % M)) % M * half % M)
        self.finish(nxt, t)

    @staticmethod
    def fold(gen_pows, code, t):  # :136-145
        f = FriProverData.init(code, t)
        for k in range(len(code).bit_length() - 1 - LOG_BLOWUP):
            f.fold_step(gen_pows, k, t.next_challenge(), t)
        assert f.last_element is not None
        return f

    def open_query_at(self, index):  # :154-174
        paths, cur, cur_n = [], index, len(self.trees[0].data)
        for m in self.trees:
            paths.append(m.open(cur))
            cur_n //= 2
            if cur_n:
                cur %= cur_n
        return paths


def _queries(t, domain_size, opener):
    qs = []
    for _ in range(NUM_QUERIES):
        idx = int.from_bytes(t.random()[:8], "little") % (domain_size // 2)
        qs.append(opener(idx))
        t.absorb(idx.to_bytes(8, "little"))
    return qs


def fri_prove(code, gen_pows, t):  # :261-285
    f = FriProverData.fold(gen_pows, code, t)
    qs = _queries(t, len(code), f.open_query_at)
    return {"commitments": [m.root() for m in f.trees], "queries": qs, "last_elem": f.last_element, "last_random": t.random()}


def _ser_fe(x):
    return (16).to_bytes(8, "little") + to_bytes(x)


def _ser_pair(v):
    return (16).to_bytes(8, "little") + v[:16] + (16).to_bytes(8, "little") + v[16:]


def _ser_path_tail(path):
    out = len(path).to_bytes(8, "little")
    for dg, d in path:
        out += dg + d.to_bytes(4, "little")
    return out


def _ser_query(paths):
    out = len(paths).to_bytes(8, "little")
    for value, path in paths:
        out += _ser_pair(value) + _ser_path_tail(path)
    return out


def fri_proof_serialize(p):  # bincode fixed-int LE via serde (src/fri/mod.rs:367-369)
    out = len(p["commitments"]).to_bytes(8, "little") + b"".join(p["commitments"])
    out += len(p["queries"]).to_bytes(8, "little")
    for q in p["queries"]:
        out += _ser_query(q)
    return out + _ser_fe(p["last_elem"]) + p["last_random"]


def query_verify(paths, commitments, last_element, n, index, gen, rs):  # :184-236
    if len(paths) != len(commitments):
        return False
    cur_n, cur_idx, cur_gen = n, index, gen
    for i, (value, path) in enumerate(paths):
        if not path_verify(value, path, commitments[i], cur_idx):
            return False
        v, mv = int.from_bytes(value[:16], "little"), int.from_bytes(value[16:], "little")
        gp = pow(cur_gen, cur_idx, M)
        even = div((v + mv) % M, 2)
        odd = div((v - mv) % M, 2 * gp % M)
        expect = (even + rs[i] * odd) % M
        if i == len(paths) - 1:
            return last_element == expect
        nxt = cur_idx % (cur_n // 2)
        nv = paths[i + 1][0]
        nv = int.from_bytes(nv[:16] if nxt == cur_idx else nv[16:], "little")
        if nv != expect:
            return False
        cur_gen, cur_n, cur_idx = cur_gen * cur_gen % M, cur_n // 2, nxt
    return True


def fri_verify_queries(p, t, rs):  # :311-340
    log_domain = len(p["commitments"]) + LOG_BLOWUP
    domain = 1 << log_domain
    gen = pow_2_generator(log_domain)
    for q in p["queries"]:
        idx = int.from_bytes(t.random()[:8], "little") % (domain // 2)
        t.absorb(idx.to_bytes(8, "little"))
        if not query_verify(q, p["commitments"], p["last_elem"], domain // 2, idx, gen, rs):
            return False
    return p["last_random"] == t.random()


def fri_verify(p):  # :287-309
    t, rs = Transcript(), []
    for c in p["commitments"]:
        t.absorb(c)
        rs.append(t.next_challenge())
    t.absorb(to_bytes(p["last_elem"]))
    return fri_verify_queries(p, t, rs)


class SumcheckTables:  # src/constraint_system/sumcheck.rs:127-277 (width 1, composition x[0])
    def __init__(self, inputs, evals):
        v = len(inputs)
        assert 1 << v == len(evals)
        self.matrix = list(evals)
        self.height = len(evals)
        self.delta = []
        for idx in range(len(evals)):  # Mask::evaluate, evaluation.rs:56-73
            p = 1
            for i in range(v):
                pt = inputs[v - 1 - i]
                p = p * (pt if (idx >> i) & 1 else (1 - pt) % M) % M
            self.delta.append(p)

    def partial_sum(self, r):  # :204-232
        off, s, acc = self.height >> 1, (1 - r) % M, 0
        for i in range(off):
            if r == 1:
                d, m = r * self.delta[i + off] % M, r * self.matrix[i + off] % M
            else:
                d = (s * self.delta[i] + r * self.delta[i + off]) % M
                m = (s * self.matrix[i] + r * self.matrix[i + off]) % M
            acc = (acc + m * d) % M
        return acc

    def fold(self, r):  # :234-247
        self.height >>= 1
        off, s = self.height, (1 - r) % M
        for i in range(off):
            self.delta[i] = (s * self.delta[i] + r * self.delta[i + off]) % M
            self.matrix[i] = (s * self.matrix[i] + r * self.matrix[i + off]) % M

    def compute_sumcheck_polynomial(self, total_degree, previous_sum, t):  # :174-202
        evals = [0] * (total_degree + 1)
        for i in range(1, total_degree + 1):
            evals[i] = self.partial_sum(from_i64(i))
        evals[0] = (previous_sum - evals[1]) % M
        pol = interpolate(evals)
        for c in pol[1:]:
            t.absorb(to_bytes(c))
        r = t.next_challenge()
        new_sum = poly_eval(pol, r)
        self.fold(r)
        return pol[1:], r, new_sum


def delta_evaluate(data, points):  # evaluation.rs:80-90
    p = 1
    for a, b in zip(data, points):
        p = p * ((a * b + (1 - a) * (1 - b)) % M) % M
    return p


def to_polynomial(nonzero, s):  # sumcheck.rs:269-276
    return [div((s - sum(nonzero)) % M, 2)] + list(nonzero)


def encode_poly(evals, gen):  # multilinear_pcs.rs:101-107
    c = to_coefficient(evals)
    bit_reverse_permutation(c)
    return reed_solomon(c, gen)


def pcs_prove(inputs, output, evals, t):  # src/fri/multilinear_pcs.rs:90-136
    log_domain = len(evals).bit_length() - 1 + LOG_BLOWUP
    gen_pows = pow_2_generator_powers(log_domain)
    code = encode_poly(evals, gen_pows[1])
    f = FriProverData.init(code, t)
    sc = SumcheckTables(inputs, evals)
    prev, polys, rs = output, [], []
    for k in range(log_domain - LOG_BLOWUP):
        nz, r, prev = sc.compute_sumcheck_polynomial(2, prev, t)
        polys.append(nz)
        rs.append(r)
        f.fold_step(gen_pows, k, r, t)
    qs = _queries(t, 1 << log_domain, f.open_query_at)
    fri = {"commitments": [m.root() for m in f.trees], "queries": qs, "last_elem": f.last_element, "last_random": t.random()}
    return {"fri": fri, "sumcheck": polys, "inputs": list(inputs), "output": output, "challenges": rs}


def _sumcheck_replay(polys, s, inputs, rs, last_elem):
    pol = to_polynomial(polys[0], s)
    for i in range(1, len(polys)):
        pol = to_polynomial(polys[i], poly_eval(pol, rs[i - 1]))
    return delta_evaluate(inputs, rs) * last_elem % M == poly_eval(pol, rs[-1])


def pcs_verify(p, t):  # :138-190
    fri, rs = p["fri"], []
    for root, nz in zip(fri["commitments"], p["sumcheck"]):
        t.absorb(root)
        for c in nz:
            t.absorb(to_bytes(c))
        rs.append(t.next_challenge())
    t.absorb(to_bytes(fri["last_elem"]))
    if not _sumcheck_replay(p["sumcheck"], p["output"], p["inputs"], rs, fri["last_elem"]):
        return False
    return fri_verify_queries(fri, t, rs)


def fingerprint(r, coeffs):  # batched_fri.rs:30-38
    acc = 0
    for c in coeffs:
        acc = (acc * r + c) % M
    return acc


def _batched_core(codes, gen_pows, t, sumcheck=None):
    """BatchedFriProverData::{init, fold} (batched_fri.rs:41-205) with the optional sumcheck
    interleave of BatchedPCSProverData::fold (batched_pcs.rs:79-127)."""
    n = len(codes[0])
    h = n // 2
    batch = Merkle.batch_commit([[to_bytes(c[i]) + to_bytes(c[i + h]) for i in range(h)] for c in codes])
    t.absorb(batch.root())
    fr = t.next_challenge()
    t.absorb(to_bytes(fr))
    f = FriProverData()
    polys, rs = [], []
    prev = None
    if sumcheck is not None:
        sc, outputs = sumcheck(fr)
        prev = fingerprint(fr, outputs)
    half = div(1, 2)
    for k in range(n.bit_length() - 1 - LOG_BLOWUP):
        if sumcheck is not None:
            nz, r, prev = sc.compute_sumcheck_polynomial(2, prev, t)
            polys.append(nz)
        else:
            r = t.next_challenge()
        rs.append(r)
        if k == 0:
            nxt = []
            for i in range(h):
                a = fingerprint(fr, [c[i] for c in codes])
                b = fingerprint(fr, [c[i + h] for c in codes])
                w = 1 if i == 0 else gen_pows[len(gen_pows) - i]
                nxt.append(((a + b) + r * ((a - b) * w % M)) % M * half % M)
            f.finish(nxt, t)
        else:
            f.fold_step(gen_pows, k, r, t)

    def opener(idx):
        return batch.open(idx), f.open_query_at(idx % (h // 2))

    qs = _queries(t, n, opener)
    fri = {"batch_commitment": batch.root(), "commitments": [m.root() for m in f.trees], "queries": qs,
           "last_elem": f.last_element, "last_random": t.random(), "fingerprint_r": fr}
    return fri, polys, rs


def batched_fri_prove(codes, gen_pows, t):  # batched_fri.rs:286-318
    return _batched_core(codes, gen_pows, t)[0]


def bfri_proof_serialize(p):
    out = p["batch_commitment"] + len(p["commitments"]).to_bytes(8, "little") + b"".join(p["commitments"])
    out += len(p["queries"]).to_bytes(8, "little")
    for (bvalue, bpath), paths in p["queries"]:
        nb = len(bvalue) // 32
        out += nb.to_bytes(8, "little") + b"".join(_ser_pair(bvalue[32 * j:32 * j + 32]) for j in range(nb))
        out += _ser_path_tail(bpath) + _ser_query(paths)
    return out + _ser_fe(p["last_elem"]) + p["last_random"]


def batched_pcs_prove(inputs, outputs, polys, t):  # batched_pcs.rs:130-180
    log_domain = len(polys[0]).bit_length() - 1 + LOG_BLOWUP
    gen_pows = pow_2_generator_powers(log_domain)
    codes = [encode_poly(p, gen_pows[1]) for p in polys]
    for x in inputs:
        t.absorb(to_bytes(x))
    for x in outputs:
        t.absorb(to_bytes(x))

    def sumcheck(fr):
        fp = [fingerprint(fr, [p[i] for p in polys]) for i in range(len(polys[0]))]
        return SumcheckTables(inputs, fp), outputs

    fri, sc_polys, rs = _batched_core(codes, gen_pows, t, sumcheck)
    return {"fri": fri, "sumcheck": sc_polys, "inputs": list(inputs), "outputs": list(outputs), "challenges": rs}
