def tree_flatten(tree, prefix="", is_leaf=None):
    out = []
    if isinstance(tree, dict):
        for k, v in tree.items():
            out.extend(tree_flatten(v, f"{prefix}.{k}" if prefix else str(k)))
    elif isinstance(tree, (list, tuple)):
        for i, v in enumerate(tree):
            out.extend(tree_flatten(v, f"{prefix}.{i}" if prefix else str(i)))
    else:
        out.append((prefix, tree))
    return out


def tree_unflatten(items):
    root = {}
    for name, v in items:
        parts = name.split(".")
        d = root
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return root


def tree_map(fn, tree, *rest):
    if isinstance(tree, dict):
        return {k: tree_map(fn, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(fn, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    return fn(tree, *rest)
