"""Independent Python big-integer restatement of the CKKS/RNS arithmetic (test-only).

Written from the mathematical definitions (evaluation at roots of unity, exact CRT
rounding, schoolbook negacyclic products) rather than from the oracle's code, so that
agreement between the two pins the oracle's restatement of SEAL 4.0 semantics
 on small rings.
"""
from fractions import Fraction


def is_prime(n):
    if n < 2:
        return False
    small = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37]
    for p in small:
        if n % p == 0:
            return n == p
    d, r = n - 1, 0
    while d % 2 == 0:
        d //= 2
        r += 1
    for a in small:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(r - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def seal_primes(N, bits, count):
    """CoeffModulus::Create order: scan down from 2^bits, assign from the back."""
    f = 2 * N
    v = ((1 << bits) - 1) // f * f + 1
    found = []
    while len(found) < count:
        if is_prime(v):
            found.append(v)
        v -= f
    return found[::-1]


def minimal_root(N, q):
    """Smallest primitive 2N-th root of unity mod q."""
    e = (q - 1) // (2 * N)
    x = 2
    while True:
        g = pow(x, e, q)
        if pow(g, N, q) == q - 1:
            break
        x += 1
    best, cur, g2 = g, g, g * g % q
    for _ in range(N):
        best = min(best, cur)
        cur = cur * g2 % q
    return best


def bitrev(x, bits):
    return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0


def ntt_eval(f, psi, q):
    """Definition: out[i] = f(psi^(2*bitrev(i)+1)), O(N^2)."""
    N = len(f)
    lg = N.bit_length() - 1
    out = []
    for i in range(N):
        x = pow(psi, 2 * bitrev(i, lg) + 1, q)
        acc = 0
        for c in reversed(f):
            acc = (acc * x + c) % q
        out.append(acc)
    return out


def ntt_fast(f, psi, q):
    """Recursive O(N log N): negacyclic evaluation, bit-reversed output order."""
    N = len(f)
    lg = N.bit_length() - 1
    a = list(f)
    # standard iterative CT with psi powers in bit-reversed order
    tw = [pow(psi, bitrev(i, lg), q) for i in range(N)]
    t = N // 2
    m = 1
    while m < N:
        for i in range(m):
            w = tw[m + i]
            base = 2 * i * t
            for j in range(base, base + t):
                u, v = a[j], a[j + t] * w % q
                a[j], a[j + t] = (u + v) % q, (u - v) % q
        m *= 2
        t //= 2
    return a


def intt_eval(F, psi, q):
    """Inverse of ntt_eval by interpolation: f[k] = N^-1 sum_i F[i] x_i^-k."""
    N = len(F)
    lg = N.bit_length() - 1
    ninv = pow(N, q - 2, q)
    xs = [pow(psi, q - 1 - (2 * bitrev(i, lg) + 1), q) for i in range(N)]  # x_i^-1
    out = []
    for k in range(N):
        acc = 0
        for i in range(N):
            acc += F[i] * pow(xs[i], k, q)
        out.append(acc % q * ninv % q)
    return out


def negacyclic_mul(a, b, q):
    N = len(a)
    r = [0] * N
    for i, x in enumerate(a):
        if x == 0:
            continue
        for j, y in enumerate(b):
            k = i + j
            if k < N:
                r[k] = (r[k] + x * y) % q
            else:
                r[k - N] = (r[k - N] - x * y) % q
    return r


def crt(residues, primes):
    Q = 1
    for p in primes:
        Q *= p
    x = 0
    for r, p in zip(residues, primes):
        M = Q // p
        x += r * M * pow(M, p - 2, p)
    return x % Q, Q


def galois_elt(step, N):
    m = 2 * N
    if step == 0:
        return m - 1
    s = step if step > 0 else N // 2 - (-step)
    return pow(3, s, m)


def galois_table(elt, N):
    lg = N.bit_length() - 1
    return [bitrev(((elt * (2 * bitrev(i, lg) + 1)) >> 1) & (N - 1), lg) for i in range(N)]


def naf(v):
    res, sign, v, i = [], v < 0, abs(v), 0
    while v:
        z = (2 - (v & 3)) if v & 1 else 0
        v = (v - z) >> 1
        if z:
            res.append((-z if sign else z) * (1 << i))
        i += 1
    return res


def rescale_poly(limbs_ntt, primes, psis):
    """Exact divide-and-round of one polynomial by its last prime (SURVEY A.2.6 / A.2.11 iv).

    limbs_ntt: list of NTT-form limbs.  Returns NTT-form limbs for primes[:-1]."""
    N = len(limbs_ntt[0])
    ql = primes[-1]
    half = ql // 2
    coeff = [intt_fast(l, psi, p) for l, psi, p in zip(limbs_ntt, psis, primes)]
    out = [[0] * N for _ in primes[:-1]]
    for k in range(N):
        x, Q = crt([c[k] for c in coeff], primes)
        y = (x + half) // ql  # round-half-up of x/ql
        for i, p in enumerate(primes[:-1]):
            out[i][k] = y % p
    return [ntt_fast(o, psi, p) for o, psi, p in zip(out, psis, primes)]


def intt_fast(F, psi, q):
    N = len(F)
    lg = N.bit_length() - 1
    a = list(F)
    ipsi = pow(psi, q - 2, q)
    tw = [pow(ipsi, bitrev(i, lg), q) for i in range(N)]
    t, m = 1, N // 2
    while m >= 1:
        for i in range(m):
            w = tw[m + i]
            base = 2 * i * t
            for j in range(base, base + t):
                u, v = a[j], a[j + t]
                a[j], a[j + t] = (u + v) % q, (u - v) * w % q
        t *= 2
        m //= 2
    ninv = pow(N, q - 2, q)
    return [x * ninv % q for x in a]


def switch_key(target_ntt, key, primes, psis, level):
    """Math-level SEAL switch_key.  target_ntt: [level][N];
    key[J][K][I] NTT-form limbs over ALL primes; returns delta[K][i] to add to the ciphertext."""
    L = len(primes)
    sp = L - 1
    N = len(target_ntt[0])
    t = [intt_fast(target_ntt[j], psis[j], primes[j]) for j in range(level)]
    idx = list(range(level)) + [sp]
    acc = [[None] * len(idx) for _ in range(2)]
    for ii, I in enumerate(idx):
        qI = primes[I]
        dig = [ntt_fast([c % qI for c in t[J]], psis[I], qI) for J in range(level)]
        for K in range(2):
            acc[K][ii] = [sum(dig[J][k] * key[J][K][I][k] for J in range(level)) % qI for k in range(N)]
    p = primes[sp]
    half = p // 2
    delta = [[None] * level for _ in range(2)]
    for K in range(2):
        r = intt_fast(acc[K][level], psis[sp], p)
        r = [(x + half) % p for x in r]
        for i in range(level):
            qi = primes[i]
            u = ntt_fast([(x - half) % qi for x in r], psis[i], qi)
            pinv = pow(p % qi, qi - 2, qi)
            delta[K][i] = [(acc[K][i][k] - u[k]) * pinv % qi for k in range(N)]
    return delta
