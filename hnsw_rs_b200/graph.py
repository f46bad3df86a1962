"""graph crate: NodeID, Dist, GraphError, Graph, Layers (host-side data types).

The device reads adjacency as fixed-stride rows (csrc/hostgraph.h); these classes are the
reference-shaped view of the same data: what HNSW.get_layer() returns, what save/load
serialise (graph/src/graph.rs:165-252) and what the reference's graph tests exercise.
"""
import struct
from functools import total_ordering

import numpy as np

NO_ID = 0xFFFFFFFF  # NodeID::MAX padding in the layer files


@total_ordering
class Dist:
    """graph/src/dist.rs:4-37: ordered by dist, ties by id; equal only if both are equal."""
    __slots__ = ("id", "dist")

    def __init__(self, id, dist):
        self.id = int(id)
        self.dist = np.float32(dist)

    def key(self):
        """u64 key whose integer order is Dist::cmp (non-negative f32 bits order like uints)."""
        return (int(np.float32(self.dist).view(np.uint32)) << 32) | self.id

    def __eq__(self, o):
        return self.id == o.id and self.dist == o.dist

    def __lt__(self, o):
        if np.isnan(self.dist) or np.isnan(o.dist):
            raise ValueError("partial_cmp().unwrap() on NaN")  # dist.rs:32 panics
        if self.dist < o.dist:
            return True
        if self.dist > o.dist:
            return False
        return self.id < o.id

    def __hash__(self):
        return hash((self.id, float(self.dist)))

    def __repr__(self):
        return f"Dist {{ id: {self.id}, dist: {self.dist} }}"


class GraphError(Exception):
    """graph/src/errors.rs:3-9"""

    def __init__(self, kind, node):
        super().__init__(f"{kind}({node})")
        self.kind = kind
        self.node = node

    NodeNotInGraph = "NodeNotInGraph"
    IsolatedNode = "IsolatedNode"
    SelfConnection = "SelfConnection"
    MExceeded = "MExceeded"


class Graph:
    """graph/src/graph.rs:11-163"""

    def __init__(self, level, m):
        self.nodes = {}  # NodeID -> set of NodeID
        self.level = level
        self.m = m

    @staticmethod
    def from_csr(level, m, node_ids, offsets, nbrs):
        g = Graph(level, m)
        for r, nid in enumerate(node_ids):
            g.nodes[int(nid)] = set(int(x) for x in nbrs[int(offsets[r]):int(offsets[r + 1])])
        return g

    def iter_nodes(self):
        return iter(self.nodes.keys())

    def add_node(self, point_id):
        self.nodes.setdefault(int(point_id), set())

    def _pair(self, a, b):
        if a in self.nodes and b in self.nodes:
            return self.nodes[a], self.nodes[b]
        if a in self.nodes:
            raise GraphError(GraphError.NodeNotInGraph, b)
        raise GraphError(GraphError.NodeNotInGraph, a)

    def add_edge(self, a, b):  # graph.rs:37-52
        if a == b:
            raise GraphError(GraphError.SelfConnection, a)
        na, nb = self._pair(a, b)
        na.add(b)
        nb.add(a)

    def remove_edge(self, a, b):  # graph.rs:72-83
        na, nb = self._pair(a, b)
        na.discard(b)
        nb.discard(a)

    def isolate_node(self, node):  # graph.rs:85-94
        for n in self.neighbors_vec(node):
            if self.degree(n) == 1:
                continue
            self.remove_edge(node, n)

    def neighbors(self, node):
        if node not in self.nodes:
            raise GraphError(GraphError.NodeNotInGraph, node)
        return set(self.nodes[node])

    def neighbors_vec(self, node):
        return list(self.neighbors(node))

    def replace_neighbors(self, node, new_neighbors):  # graph.rs:128-137
        self.isolate_node(node)
        self.add_neighbors(node, new_neighbors)

    def add_neighbors(self, node, new_neighbors):
        for n in new_neighbors:
            self.add_edge(node, n)

    def degree(self, node):
        if node not in self.nodes:
            raise GraphError(GraphError.NodeNotInGraph, node)
        return len(self.nodes[node])

    def nb_nodes(self):
        return len(self.nodes)

    def contains(self, node):
        return node in self.nodes

    # Serializer, graph.rs:201-252 (row width = m; see SURVEY App. B hazard 1)
    def size(self):
        return 7 + self.nb_nodes() * 4 * (self.m + 1)

    def serialize(self):
        out = [struct.pack(">B", self.level), struct.pack(">I", self.nb_nodes()), struct.pack(">H", self.m)]
        for node in self.iter_nodes():
            nb = list(self.nodes[node])
            nb += [NO_ID] * max(0, self.m - len(nb))
            out.append(struct.pack(f">{1 + len(nb)}I", node, *nb))
        return b"".join(out)

    @staticmethod
    def deserialize(data):
        level = data[0]
        (nb_nodes,) = struct.unpack(">I", data[1:5])
        (m,) = struct.unpack(">H", data[5:7])
        g = Graph(level, m)
        i = 7
        for _ in range(nb_nodes):
            row = struct.unpack(f">{1 + m}I", data[i:i + 4 * (1 + m)])
            i += 4 * (1 + m)
            nbrs = set()
            for v in row[1:]:
                if v == NO_ID:
                    break
                nbrs.add(v)
            g.nodes[row[0]] = nbrs
        return g


class Layers:
    """graph/src/layers.rs:7-71"""

    def __init__(self, m):
        self.levels = []
        self.m = m

    def len(self):
        return len(self.levels)

    __len__ = len

    def get_layer(self, layer_nb):
        if layer_nb >= len(self.levels):
            raise IndexError(f"Layer {layer_nb} not found in the structure.")  # layers.rs:28 panics
        return self.levels[layer_nb]

    get_layer_mut = get_layer

    def add_layer(self, graph):
        self.levels.append(graph)

    def iter_layers(self):
        return iter(self.levels)

    def add_level(self, level):  # layers.rs:48-60: layer 0 cap 2m, others m
        while len(self.levels) <= level:
            m = self.m * 2 if len(self.levels) == 0 else self.m
            self.add_layer(Graph(len(self.levels), m))

    def add_node(self, point_id, level):  # layers.rs:62-70
        self.add_level(level)
        for layer in self.levels[:level + 1]:
            layer.add_node(point_id)
