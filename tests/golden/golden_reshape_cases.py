"""Random reshape requests (merges, splits, kept axes, length-1 axes) shared by generate_reshape.py and
tests/test_reshape_plan.py.  TEST INFRASTRUCTURE ONLY."""


def _factors(rng, n, k):
    fs, m, p, primes = [1] * k, n, 2, []
    while m > 1:
        while m % p == 0:
            primes.append(p)
            m //= p
        p += 1
    for q in primes:
        fs[rng.randrange(k)] *= q
    return fs


def _blocks(rng, n):
    if rng.random() < 0.5:
        c = min(n, rng.choice([1, 2, 3, 4, 5, 8, n]))
        full, rest = divmod(n, c)
        return (c,) * full + ((rest,) if rest else ())
    parts, rem = [], n
    while rem > 0:
        t = rng.randint(1, max(1, min(rem, rng.choice([2, 4, 9]))))
        parts.append(t)
        rem -= t
    return tuple(parts)


def random_case(rng):
    inshape, outshape = [], []
    for _ in range(rng.randint(1, 3)):
        n = rng.choice([1, 2, 4, 6, 8, 12, 16, 24, 30, 36, 60])
        mode = rng.random()
        if mode < 0.3:
            inshape.append(n); outshape.append(n)
        elif mode < 0.6:
            inshape += _factors(rng, n, rng.randint(2, 3)); outshape.append(n)
        elif mode < 0.9:
            inshape.append(n); outshape += _factors(rng, n, rng.randint(2, 3))
        else:
            inshape += _factors(rng, n, 2); outshape += _factors(rng, n, 2)
        if rng.random() < 0.15:
            inshape.insert(rng.randrange(len(inshape) + 1), 1)
        if rng.random() < 0.15:
            outshape.insert(rng.randrange(len(outshape) + 1), 1)
    inshape, outshape = tuple(inshape), tuple(outshape)
    return inshape, outshape, tuple(_blocks(rng, n) for n in inshape)
