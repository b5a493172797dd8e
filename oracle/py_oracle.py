"""ORACLE (test infrastructure only) — pure-Python restatement for tiny graphs.

Follows /root/reference src/generic_graph/generic_graph.rs:1310-1385 line by line
(queues, `ordering` stack, predecessor lists, `b += b_k; b -= 1.0`).  Used to
cross-check the C oracle (oracle/ref_vertex_load.c) on small inputs; never imported
by the product package.
"""
from collections import deque


def vertex_load(adj, include_endpoints):
    """adj: list of neighbour lists in storage order.  Returns list[float]."""
    n = len(adj)
    b = [0.0] * n
    for i in range(n):                                   # :1321
        b_k = [1.0] * n                                  # :1327
        distance = [None] * n                            # :1328
        predecessor = [[] for _ in range(n)]             # :1330
        ordering = []
        depth = 0                                        # :1335
        queue0, queue1 = deque([i]), deque()             # :1336
        distance[i] = depth                              # :1337
        while queue0:                                    # :1341
            index = queue0.popleft()
            ordering.append(index)                       # :1342
            for nb in adj[index]:                        # :1344
                d = distance[nb]
                if d is not None:                        # :1345
                    if d == depth + 1:
                        predecessor[nb].append(index)    # :1347
                else:
                    distance[nb] = depth + 1             # :1352
                    queue1.append(nb)
                    predecessor[nb].append(index)        # :1354
            if not queue0:                               # :1357
                queue0, queue1 = queue1, queue0
                depth += 1
        while ordering:                                  # :1364
            index = ordering.pop()
            if not ordering:                             # :1366
                break
            b[index] += b_k[index]                       # :1371
            if not include_endpoints:
                b[index] -= 1.0                          # :1373
            fraction = b_k[index] / float(len(predecessor[index]))  # :1377
            for p in predecessor[index]:
                b_k[p] += fraction                       # :1379
    return b
