"""Rooted tree with the slice of the ete3 ``Tree`` behaviour the PT test relies on.

The reference builds its tree with the third-party package ete3 (unpinned in
its README; behaviour restated here from ete3 3.1.x's TreeNode: Newick formats 1
and 2, ``set_outgroup``, ``delete`` with ``prevent_nondicotomic``, level-order
``iter_descendants`` and pre-order leaf iteration).  The order in which children
are kept and nodes are visited defines the row order of the clade matrix
(run_PT_test.py:391-394) and the branch order of the LRT model
(run_PT_test.py:569-622), so it is reproduced exactly, not just the topology.
"""
from collections import deque
import re


class NewickError(ValueError):
    pass


class TreeError(ValueError):
    pass


_FLOAT = r"[+-]?(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?"


class Tree:
    def __init__(self, newick=None, format=0, dist=None, support=None, name=None):
        self.children = []
        self.up = None
        self.dist = 1.0
        self.support = 1.0
        self.name = name if name else ""
        self.features = {"dist", "support", "name"}
        if dist is not None:
            self.dist = dist
        if support is not None:
            self.support = support
        if newick is not None:
            self.dist = 0.0
            _parse_newick(newick, self, format)

    # ------------------------------------------------------------- structure
    def add_child(self, child=None, name=None, dist=None, support=None):
        if child is None:
            child = self.__class__()
        if name is not None:
            child.name = name
        if dist is not None:
            child.dist = dist
        if support is not None:
            child.support = support
        self.children.append(child)
        child.up = self
        return child

    def remove_child(self, child):
        try:
            self.children.remove(child)
        except ValueError:
            raise TreeError("child not found")
        child.up = None
        return child

    def delete(self, prevent_nondicotomic=True, preserve_branch_length=False):
        """Remove this node and hand its children to its parent; a parent left with
        a single child is collapsed too (its child is appended to the grandparent)."""
        parent = self.up
        if parent:
            if preserve_branch_length:
                if len(self.children) == 1:
                    self.children[0].dist += self.dist
                elif len(self.children) > 1:
                    parent.dist += self.dist
            for ch in list(self.children):
                parent.add_child(ch)
            parent.remove_child(self)
        if prevent_nondicotomic and parent and len(parent.children) < 2:
            parent.delete(prevent_nondicotomic=False, preserve_branch_length=preserve_branch_length)

    def is_leaf(self):
        return len(self.children) == 0

    def is_root(self):
        return self.up is None

    def get_tree_root(self):
        node = self
        while node.up is not None:
            node = node.up
        return node

    def get_ancestors(self):
        out = []
        node = self.up
        while node is not None:
            out.append(node)
            node = node.up
        return out

    def add_features(self, **kwargs):
        for key, value in kwargs.items():
            setattr(self, key, value)
            self.features.add(key)

    def add_feature(self, key, value):
        setattr(self, key, value)
        self.features.add(key)

    # ------------------------------------------------------------- traversal
    def traverse(self, strategy="levelorder"):
        if strategy == "levelorder":
            todo = deque([self])
            while todo:
                node = todo.popleft()
                yield node
                todo.extend(node.children)
        elif strategy == "preorder":
            todo = [self]
            while todo:
                node = todo.pop()
                yield node
                todo.extend(reversed(node.children))
        elif strategy == "postorder":
            todo = [(self, False)]
            while todo:
                node, done = todo.pop()
                if done or not node.children:
                    yield node
                else:
                    todo.append((node, True))
                    for ch in reversed(node.children):
                        todo.append((ch, False))
        else:
            raise TreeError("unknown strategy %r" % (strategy,))

    def iter_descendants(self, strategy="levelorder"):
        for node in self.traverse(strategy):
            if node is not self:
                yield node

    def get_descendants(self, strategy="levelorder"):
        return list(self.iter_descendants(strategy))

    def iter_leaves(self):
        for node in self.traverse("preorder"):
            if not node.children:
                yield node

    def get_leaves(self):
        return list(self.iter_leaves())

    def get_leaf_names(self):
        return [n.name for n in self.iter_leaves()]

    def __len__(self):
        return sum(1 for _ in self.iter_leaves())

    def __iter__(self):
        return self.iter_leaves()

    def search_nodes(self, **conditions):
        out = []
        for node in self.traverse():
            if all(hasattr(node, k) and getattr(node, k) == v for k, v in conditions.items()):
                out.append(node)
        return out

    def __and__(self, value):
        value = str(value)
        for node in self.traverse():
            if node.name == value:
                return node
        raise TreeError("Node not found")

    # --------------------------------------------------------------- rooting
    def set_outgroup(self, outgroup):
        """Re-root on the branch above ``outgroup`` (which becomes the first child)."""
        if isinstance(outgroup, str):
            outgroup = self & outgroup
        if self is outgroup:
            raise TreeError("Cannot set myself as outgroup")
        parent_outgroup = outgroup.up
        # child of self on the path to the outgroup
        n = outgroup
        while n.up is not self:
            n = n.up
        self.children.remove(n)
        if len(self.children) != 1:
            connector = self.__class__()
            connector.dist = 0.0
            connector.support = n.support
            for ch in list(self.children):
                connector.children.append(ch)
                ch.up = connector
                self.children.remove(ch)
        else:
            connector = self.children[0]
        new_parent = parent_outgroup
        if new_parent is not self:
            # reverse the parent/child direction along the path self -> parent_outgroup
            new_child = new_parent.up
            previous = None
            buf_dist, buf_support = new_parent.dist, new_parent.support
            while new_child is not self:
                new_parent.children.append(new_child)
                new_child.children.remove(new_parent)
                nxt_dist, nxt_support = new_child.dist, new_child.support
                new_child.dist, new_child.support = buf_dist, buf_support
                buf_dist, buf_support = nxt_dist, nxt_support
                new_parent.up = previous
                previous = new_parent
                new_parent = new_child
                new_child = new_parent.up
            new_parent.children.append(connector)
            connector.up = new_parent
            new_parent.up = previous
            connector.dist += buf_dist
            outgroup2 = parent_outgroup
            parent_outgroup.children.remove(outgroup)
            outgroup2.dist = 0
        else:
            outgroup2 = connector
        outgroup.up = self
        outgroup2.up = self
        self.children = [outgroup, outgroup2]
        middist = (outgroup2.dist + outgroup.dist) / 2
        outgroup.dist = middist
        outgroup2.dist = middist
        outgroup2.support = outgroup.support

    # ---------------------------------------------------------------- output
    def write(self):
        def rec(node):
            if node.children:
                return "(" + ",".join(rec(c) for c in node.children) + ")%s:%g" % (node.name if False else "", node.dist)
            return "%s:%g" % (node.name, node.dist)

        if self.children:
            return "(" + ",".join(rec(c) for c in self.children) + ");"
        return self.name + ";"


TreeNode = Tree


# ------------------------------------------------------------------- Newick
def _parse_newick(text, root, fmt):
    """Formats 1 (flexible, internal node names) and 2 (strict: every internal
    node carries a support value and a length, every leaf a name and a length)."""
    if fmt not in (0, 1, 2):
        raise NewickError("unsupported newick format %r" % (fmt,))
    text = re.sub(r"[\n\r\t]+", "", text.strip())
    if not text.endswith(";"):
        raise NewickError("Unexisting tree file or Malformed newick tree structure.")
    if text.count("(") != text.count(")"):
        raise NewickError("Parentheses do not match. Broken tree structure?")
    n = len(text)
    if not text.startswith("("):
        _read_internal(text[:-1], root, fmt)
        return
    i = 0
    cur = None
    expect_item = False
    while i < n:
        ch = text[i]
        if ch == "(":
            cur = root if cur is None else cur.add_child()
            i += 1
            expect_item = True
        elif ch == ",":
            if expect_item:
                raise NewickError("Empty leaf node found")
            i += 1
            expect_item = True
        elif ch == ")":
            if expect_item:
                raise NewickError("Empty leaf node found")
            j = i + 1
            while j < n and text[j] not in ",();":
                j += 1
            _read_internal(text[i + 1:j], cur, fmt)
            i = j
            if cur is root:
                break
            cur = cur.up
        elif ch == ";":
            break
        else:
            if cur is None:
                raise NewickError("Unexisting tree file or Malformed newick tree structure.")
            i = _read_leaf(text, i, cur, fmt)
            expect_item = False


def _read_leaf(text, i, parent, fmt):
    j = i
    n = len(text)
    while j < n and text[j] not in ",();":
        j += 1
    label = text[i:j].strip()
    if not label:
        raise NewickError("Empty leaf node found")
    leaf = parent.add_child()
    name, dist = _split_label(label)
    if fmt == 2 and (name == "" or dist is None):
        raise NewickError("Unexpected newick format '%s'" % label[:50])
    leaf.name = name
    if dist is not None:
        leaf.dist = dist
    return j


def _read_internal(label, node, fmt):
    label = label.strip()
    if not label:
        return
    first, dist = _split_label(label)
    if fmt == 2:
        if first == "" or dist is None or not re.fullmatch(_FLOAT, first):
            raise NewickError("Unexpected newick format '%s'" % label[:50])
        node.support = float(first)
    elif fmt == 0:
        if first != "":
            if not re.fullmatch(_FLOAT, first):
                raise NewickError("Unexpected newick format '%s'" % label[:50])
            node.support = float(first)
    else:
        node.name = first
    if dist is not None:
        node.dist = dist


def _split_label(label):
    label = re.sub(r"\[&&NHX:[^\]]*\]$", "", label)
    if ":" in label:
        name, _, d = label.rpartition(":")
        d = d.strip()
        if not re.fullmatch(_FLOAT, d):
            raise NewickError("Unexpected newick format '%s'" % label[:50])
        return name.strip(), float(d)
    return label.strip(), None
