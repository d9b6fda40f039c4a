"""Test helper: write a list of reference-style elements as TU-format text files
(<name>_graph_indicator.txt, _A.txt, _node_labels.txt, _edge_labels.txt, _graph_labels.txt,
_node_attributes.txt) -- the on-disk format grakel.datasets.read_data parses (datasets/base.py:135-290).
Node ids must be the global 1-based ids of the files (as in the bundled MUTAG fixture)."""
import os


def write_tu(path, name, elements, classes=None, edge_labels=None, attributes=None, node_labels=True):
    """elements: [[iterable of (u, v), {node id: int label}], ...] with global 1-based node ids, graphs in
    file order and node ids increasing from graph to graph.  `edge_labels`: one {(u, v): int} per graph."""
    os.makedirs(path, exist_ok=True)
    base = os.path.join(path, name + "_")
    nodes, edges = [], []
    for gi, el in enumerate(elements, 1):
        for v in sorted(el[1]):
            nodes.append((v, gi, el[1][v]))
        for (a, b) in el[0]:
            edges.append((a, b, None if edge_labels is None else edge_labels[gi - 1][(a, b)]))
    assert [n[0] for n in nodes] == list(range(1, len(nodes) + 1)), "node ids must be 1..n in graph order"
    with open(base + "graph_indicator.txt", "w") as f:
        f.write("".join(f"{g}\n" for _, g, _ in nodes))
    with open(base + "A.txt", "w") as f:
        f.write("".join(f"{a}, {b}\n" for a, b, _ in edges))
    if node_labels:
        with open(base + "node_labels.txt", "w") as f:
            f.write("".join(f"{l}\n" for _, _, l in nodes))
    if edge_labels is not None:
        with open(base + "edge_labels.txt", "w") as f:
            f.write("".join(f"{l}\n" for _, _, l in edges))
    if classes is not None:
        with open(base + "graph_labels.txt", "w") as f:
            f.write("".join(f"{int(c)}\n" for c in classes))
    if attributes is not None:  # {node id: sequence of floats}
        with open(base + "node_attributes.txt", "w") as f:
            f.write("".join(", ".join(repr(float(x)) for x in attributes[v]) + "\n" for v, _, _ in nodes))


def renumber(elements):
    """Elements with per-graph vertex ids 0..n-1 (the seeded generators) -> global 1-based node ids."""
    out, base = [], 1
    for g, l in elements:
        out.append([[(a + base, b + base) for (a, b) in g], {v + base: lab for v, lab in l.items()}])
        base += len(l)
    return out
