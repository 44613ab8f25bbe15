def product(xs):
    out = 1
    for x in xs:
        out *= x
    return out
