// scs_tree.cu -- host-side tree work of a replicate sweep, in native code and on host threads.
//
// A simulation sweep (BASELINE configs[4]) runs the PT test on thousands of small (VCF, tree) pairs.  The device side of
// a pair costs a few microseconds; read_tree (run_PT_test.py:210-334) and the clade rows of map_mutations_gt (:389-397)
// in Python cost milliseconds (regexes, ete3-style tree objects, one Python loop per node).  This file restates exactly
// that host work for CellPhy / plain Newick trees:
//     text clean-up and the cell -> tumcell / outgcell -> healthycell renames (:217-221, :282-285),
//     Newick parsing (ete3 formats 1 / 2: only topology and leaf names matter to the PT test),
//     tree & 'healthycell', set_outgroup, delete (:301-308), deletion of leaves that are not VCF samples (:311-313),
//     tree.iter_descendants() in level order, and the membership mask of every node (:391-394),
// with the child order of every step preserved -- it defines the row order of the clade matrix and, later, the branch
// order of the LRT model.  The reference uses ete3; the semantics restated here are those of ete3 3.1.x TreeNode as
// scsommerclock_b200/tree.py documents them, and tests/test_host_logic.py plays the two against each other.
// infSCITE trees (:233-280) keep the Python path.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <initializer_list>
#include <unordered_map>
#include <vector>

#include "scs_internal.h"

namespace scs {
namespace {

// children of a node: four inline slots (a binary tree never needs more than the three a node holds for a moment while one
// of its children is being deleted), a heap vector beyond that (multifurcating input) -- 200 nodes no longer cost 200 allocations
struct ChildList {
    int inl[4];
    int n = 0;
    std::vector<int>* big = nullptr;
    ChildList() = default;
    ChildList(const ChildList& o) : n(o.n), big(o.big ? new std::vector<int>(*o.big) : nullptr) { memcpy(inl, o.inl, sizeof(inl)); }
    ChildList(ChildList&& o) noexcept : n(o.n), big(o.big) {
        memcpy(inl, o.inl, sizeof(inl));
        o.big = nullptr;
        o.n = 0;
    }
    ChildList& operator=(const ChildList& o) {
        if (this != &o) {
            delete big;
            n = o.n;
            big = o.big ? new std::vector<int>(*o.big) : nullptr;
            memcpy(inl, o.inl, sizeof(inl));
        }
        return *this;
    }
    ChildList& operator=(std::initializer_list<int> l) {
        clear();
        for (int v : l) push_back(v);
        return *this;
    }
    ~ChildList() { delete big; }
    int* begin() { return big ? big->data() : inl; }
    int* end() { return begin() + n; }
    const int* begin() const { return big ? big->data() : inl; }
    const int* end() const { return begin() + n; }
    size_t size() const { return size_t(n); }
    bool empty() const { return n == 0; }
    int operator[](size_t i) const { return begin()[i]; }
    void push_back(int v) {
        if (!big && n == 4) big = new std::vector<int>(inl, inl + 4);
        if (big) big->push_back(v);
        else inl[n] = v;
        ++n;
    }
    void clear() {
        delete big;
        big = nullptr;
        n = 0;
    }
    void erase(int* it) {
        if (big) {
            big->erase(big->begin() + (it - big->data()));
        } else {
            for (int* p = it; p + 1 < inl + n; ++p) *p = p[1];
        }
        --n;
    }
};

struct TNode {
    int up = -1;
    ChildList ch;
    int name_lo = 0, name_len = 0;      // leaf / internal label (without the distance) inside the cleaned text
    int col = -1;                       // a leaf's sample column once it has been looked up
};

struct TreeBuilder {
    std::vector<TNode> nd;
    std::string text;
    std::string err;

    int new_node() {
        nd.emplace_back();
        return int(nd.size()) - 1;
    }
    void add_child(int parent, int child) {
        nd[parent].ch.push_back(child);
        nd[child].up = parent;
    }
    void remove_child(int parent, int child) {
        auto& v = nd[parent].ch;
        v.erase(std::find(v.begin(), v.end(), child));
        nd[child].up = -1;
    }
    // TreeNode.delete (tree.py:64-78): the node's children go to its parent (appended), a parent left with a single
    // child is removed the same way (its child is appended to the grandparent)
    void del(int node, bool prevent_nondicotomic) {
        const int parent = nd[node].up;
        if (parent >= 0) {
            const ChildList kids = nd[node].ch;
            for (int c : kids) add_child(parent, c);
            nd[node].ch.clear();
            remove_child(parent, node);
        }
        if (prevent_nondicotomic && parent >= 0 && nd[parent].ch.size() < 2) del(parent, false);
    }
    bool name_is(int node, const char* s, size_t n) const {
        return size_t(nd[node].name_len) == n && memcmp(text.data() + nd[node].name_lo, s, n) == 0;
    }
};

bool is_float(const char* s, size_t n) {
    // [+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?
    size_t i = 0;
    if (i < n && (s[i] == '+' || s[i] == '-')) ++i;
    size_t d0 = i;
    while (i < n && s[i] >= '0' && s[i] <= '9') ++i;
    size_t nd1 = i - d0, nd2 = 0;
    if (i < n && s[i] == '.') {
        ++i;
        size_t f0 = i;
        while (i < n && s[i] >= '0' && s[i] <= '9') ++i;
        nd2 = i - f0;
    }
    if (nd1 == 0 && nd2 == 0) return false;
    if (i < n && (s[i] == 'e' || s[i] == 'E')) {
        ++i;
        if (i < n && (s[i] == '+' || s[i] == '-')) ++i;
        size_t e0 = i;
        while (i < n && s[i] >= '0' && s[i] <= '9') ++i;
        if (i == e0) return false;
    }
    return i == n;
}

// label = name[:dist] (split at the LAST colon, tree.py:_split_label); false when the distance is not a number
bool split_label(const std::string& t, int lo, int hi, int* name_lo, int* name_len) {
    while (lo < hi && (t[lo] == ' ')) ++lo;
    while (hi > lo && (t[hi - 1] == ' ')) --hi;
    // a trailing [&&NHX:...] comment is dropped
    if (hi > lo && t[hi - 1] == ']') {
        const size_t p = t.rfind("[&&NHX:", size_t(hi - 1));
        if (p != std::string::npos && int(p) >= lo) hi = int(p);
    }
    int colon = -1;
    for (int i = hi - 1; i >= lo; --i)
        if (t[i] == ':') {
            colon = i;
            break;
        }
    int nlo = lo, nhi = hi;
    if (colon >= 0) {
        int dlo = colon + 1, dhi = hi;
        while (dlo < dhi && t[dlo] == ' ') ++dlo;
        while (dhi > dlo && t[dhi - 1] == ' ') --dhi;
        if (!is_float(t.data() + dlo, size_t(dhi - dlo))) return false;
        nhi = colon;
    }
    while (nlo < nhi && t[nlo] == ' ') ++nlo;
    while (nhi > nlo && t[nhi - 1] == ' ') --nhi;
    *name_lo = nlo;
    *name_len = nhi - nlo;
    return true;
}

// read_tree's text clean-up for CellPhy / plain Newick files
std::string clean_text(const char* s, size_t n, bool strip_tags) {
    size_t a = 0, b = n;
    auto ws = [](char c) { return c == ' ' || c == '\n' || c == '\r' || c == '\t' || c == '\f' || c == '\v'; };
    while (a < b && ws(s[a])) ++a;
    while (b > a && ws(s[b - 1])) --b;
    std::string t(s + a, b - a);
    if (strip_tags) {                                   // re.sub('\[\d+\]', '', tree_raw)
        std::string o;
        o.reserve(t.size());
        for (size_t i = 0; i < t.size();) {
            if (t[i] == '[') {
                size_t j = i + 1;
                while (j < t.size() && t[j] >= '0' && t[j] <= '9') ++j;
                if (j > i + 1 && j < t.size() && t[j] == ']') {
                    i = j + 1;
                    continue;
                }
            }
            o.push_back(t[i++]);
        }
        t.swap(o);
    }
    if (t.find("tumcell") == std::string::npos) {       // re.sub('cell(?=\d+)', 'tumcell', tree_raw)
        std::string o;
        o.reserve(t.size() + t.size() / 4);
        for (size_t i = 0; i < t.size();) {
            if (i + 4 < t.size() && t.compare(i, 4, "cell") == 0 && t[i + 4] >= '0' && t[i + 4] <= '9') {
                o += "tumcell";
                i += 4;
            } else {
                o.push_back(t[i++]);
            }
        }
        t.swap(o);
    }
    if (t.find("healthycell") == std::string::npos) {   // re.sub('outgcell', 'healthycell', tree_raw)
        std::string o;
        o.reserve(t.size() + 16);
        for (size_t i = 0; i < t.size();) {
            if (t.compare(i, 8, "outgcell") == 0) {
                o += "healthycell";
                i += 8;
            } else {
                o.push_back(t[i++]);
            }
        }
        t.swap(o);
    }
    // the parser's own clean-up (tree.py:_parse_newick): line breaks and tabs are dropped
    if (!memchr(t.data(), '\n', t.size()) && !memchr(t.data(), '\r', t.size()) && !memchr(t.data(), '\t', t.size())) return t;
    std::string o;
    o.reserve(t.size());
    for (char c : t)
        if (c != '\n' && c != '\r' && c != '\t') o.push_back(c);
    return o;
}

// Newick -> nodes (tree.py:_parse_newick, format 1 semantics: labels of internal nodes are names; format 2 differs only
// in reading them as support values, which the PT test never looks at).  Node 0 is the root.
bool parse_newick(TreeBuilder& B) {
    const std::string& t = B.text;
    const int n = int(t.size());
    if (n == 0 || t[n - 1] != ';') {
        B.err = "Unexisting tree file or Malformed newick tree structure.";
        return false;
    }
    if (std::count(t.begin(), t.end(), '(') != std::count(t.begin(), t.end(), ')')) {
        B.err = "Parentheses do not match. Broken tree structure?";
        return false;
    }
    B.nd.reserve(size_t(2 * std::count(t.begin(), t.end(), ',') + 8));      // (a binary tree has 2 leaves - 1 nodes; + the connector of set_outgroup)
    const int root = B.new_node();
    if (t[0] != '(') {                                  // a single node
        if (!split_label(t, 0, n - 1, &B.nd[root].name_lo, &B.nd[root].name_len)) {
            B.err = "Unexpected newick format";
            return false;
        }
        return true;
    }
    int cur = -1;
    bool expect_item = false;
    int i = 0;
    while (i < n) {
        const char ch = t[i];
        if (ch == '(') {
            if (cur < 0) cur = root;
            else {
                const int c = B.new_node();
                B.add_child(cur, c);
                cur = c;
            }
            ++i;
            expect_item = true;
        } else if (ch == ',') {
            if (expect_item) {
                B.err = "Empty leaf node found";
                return false;
            }
            ++i;
            expect_item = true;
        } else if (ch == ')') {
            if (expect_item) {
                B.err = "Empty leaf node found";
                return false;
            }
            int j = i + 1;
            while (j < n && t[j] != ',' && t[j] != '(' && t[j] != ')' && t[j] != ';') ++j;
            if (cur < 0) {
                B.err = "Unexisting tree file or Malformed newick tree structure.";
                return false;
            }
            // (an empty label leaves the node untouched)
            bool blank = true;
            for (int k = i + 1; k < j; ++k)
                if (t[k] != ' ') blank = false;
            if (!blank && !split_label(t, i + 1, j, &B.nd[cur].name_lo, &B.nd[cur].name_len)) {
                B.err = "Unexpected newick format";
                return false;
            }
            i = j;
            if (cur == root) break;
            cur = B.nd[cur].up;
        } else if (ch == ';') {
            break;
        } else {
            if (cur < 0) {
                B.err = "Unexisting tree file or Malformed newick tree structure.";
                return false;
            }
            int j = i;
            while (j < n && t[j] != ',' && t[j] != '(' && t[j] != ')' && t[j] != ';') ++j;
            bool blank = true;
            for (int k = i; k < j; ++k)
                if (t[k] != ' ') blank = false;
            if (blank) {
                B.err = "Empty leaf node found";
                return false;
            }
            const int leaf = B.new_node();
            B.add_child(cur, leaf);
            if (!split_label(t, i, j, &B.nd[leaf].name_lo, &B.nd[leaf].name_len)) {
                B.err = "Unexpected newick format";
                return false;
            }
            i = j;
            expect_item = false;
        }
    }
    return true;
}

// TreeNode.set_outgroup (tree.py:176-229) without the distances: re-root on the branch above `outgroup`, which becomes
// the first child of the root
void set_outgroup(TreeBuilder& B, int root, int outgroup) {
    auto& nd = B.nd;
    const int parent_outgroup = nd[outgroup].up;
    int n = outgroup;
    while (nd[n].up != root) n = nd[n].up;
    {
        auto& v = nd[root].ch;
        v.erase(std::find(v.begin(), v.end(), n));
    }
    int connector;
    if (nd[root].ch.size() != 1) {
        connector = B.new_node();
        const ChildList kids = B.nd[root].ch;
        for (int c : kids) {
            B.nd[connector].ch.push_back(c);
            B.nd[c].up = connector;
        }
        B.nd[root].ch.clear();
    } else {
        connector = nd[root].ch[0];
    }
    int outgroup2;
    int new_parent = parent_outgroup;
    if (new_parent != root) {
        // reverse the parent / child direction along the path root -> parent_outgroup
        int new_child = B.nd[new_parent].up;
        int previous = -1;
        while (new_child != root) {
            B.nd[new_parent].ch.push_back(new_child);
            auto& v = B.nd[new_child].ch;
            v.erase(std::find(v.begin(), v.end(), new_parent));
            B.nd[new_parent].up = previous;
            previous = new_parent;
            new_parent = new_child;
            new_child = B.nd[new_parent].up;
        }
        B.nd[new_parent].ch.push_back(connector);
        B.nd[connector].up = new_parent;
        B.nd[new_parent].up = previous;
        outgroup2 = parent_outgroup;
        auto& v = B.nd[parent_outgroup].ch;
        v.erase(std::find(v.begin(), v.end(), outgroup));
    } else {
        outgroup2 = connector;
    }
    B.nd[outgroup].up = root;
    B.nd[outgroup2].up = root;
    B.nd[root].ch = {outgroup, outgroup2};
}

// sample name -> column: open addressing over FNV-1a of the name, no allocation per lookup (a leaf name used to cost a
// std::string and an unordered_map probe, twice)
struct SampleIndex {
    std::vector<int> slot;          // column or -1
    const char* names = nullptr;
    const int64_t* off = nullptr;
    uint32_t mask = 0;
    static uint32_t hash(const char* s, size_t n) {
        uint32_t h = 2166136261u;
        for (size_t i = 0; i < n; ++i) h = (h ^ uint8_t(s[i])) * 16777619u;
        return h;
    }
    void build(const char* sample_names, const int64_t* name_off, int n) {
        names = sample_names;
        off = name_off;
        uint32_t cap = 16;
        while (cap < uint32_t(4 * n)) cap <<= 1;
        mask = cap - 1;
        slot.assign(cap, -1);
        for (int c = 0; c < n; ++c) {
            uint32_t h = hash(names + off[c], size_t(off[c + 1] - off[c])) & mask;
            while (slot[h] >= 0) {
                const int o = slot[h];          // (a duplicated sample name keeps its first column, like dict insertion would not -- VCF samples are unique)
                if (off[o + 1] - off[o] == off[c + 1] - off[c] && memcmp(names + off[o], names + off[c], size_t(off[c + 1] - off[c])) == 0) break;
                h = (h + 1) & mask;
            }
            if (slot[h] < 0) slot[h] = c;
        }
    }
    int find(const char* s, size_t n) const {
        uint32_t h = hash(s, n) & mask;
        while (slot[h] >= 0) {
            const int o = slot[h];
            if (size_t(off[o + 1] - off[o]) == n && memcmp(names + off[o], s, n) == 0) return o;
            h = (h + 1) & mask;
        }
        return -1;
    }
};

// One tree: everything read_tree and the S rows need.  Returns false with B.err set on failure.
bool process_tree(const char* s, size_t n, bool strip_tags, const SampleIndex& samples, const std::string& outg, int n_rows,
                  int64_t pitch_words, uint32_t* bits, int32_t* children, int32_t* info, std::string* err) {
    TreeBuilder B;
    B.text = clean_text(s, n, strip_tags);
    if (!parse_newick(B)) {
        *err = B.err;
        return false;
    }
    const int root = 0;
    // tree & 'healthycell': first node in level order with that name
    std::vector<int> order;
    order.reserve(B.nd.size());
    order.push_back(root);
    for (size_t h = 0; h < order.size(); ++h)
        for (int c : B.nd[order[h]].ch) order.push_back(c);
    int outg_node = -1;
    for (int v : order)
        if (B.name_is(v, outg.data(), outg.size())) {
            outg_node = v;
            break;
        }
    if (outg_node < 0 || outg_node == root) {
        *err = "Node not found: the tree has no leaf named " + outg;
        return false;
    }
    set_outgroup(B, root, outg_node);
    B.del(outg_node, true);
    // leaves that are not VCF samples are removed (:311-313); the list of leaves is taken once, in pre-order
    {
        std::vector<int> leaves, stack{root};
        while (!stack.empty()) {
            const int v = stack.back();
            stack.pop_back();
            if (B.nd[v].ch.empty()) leaves.push_back(v);
            for (const int* it = B.nd[v].ch.end(); it != B.nd[v].ch.begin();) stack.push_back(*--it);
        }
        for (int v : leaves) {
            B.nd[v].col = samples.find(B.text.data() + B.nd[v].name_lo, size_t(B.nd[v].name_len));
            if (B.nd[v].col < 0) B.del(v, true);
        }
    }
    // tree.iter_descendants(): level order below the root
    order.clear();
    for (int c : B.nd[root].ch) order.push_back(c);
    for (size_t h = 0; h < order.size(); ++h)
        for (int c : B.nd[order[h]].ch) order.push_back(c);
    const int n_nodes = int(order.size());
    int n_leaves = 0;
    for (int v : order) n_leaves += B.nd[v].ch.empty();
    if (n_nodes == 0 || n_nodes != 2 * n_leaves - 1) {
        *err = "the PT test needs a binary tree: " + std::to_string(n_leaves) + " leaves but " + std::to_string(n_nodes) + " nodes";
        return false;
    }
    if (n_nodes + 1 > n_rows) {
        *err = "tree has " + std::to_string(n_nodes) + " nodes, more than the " + std::to_string(n_rows - 1) + " clade rows planned";
        return false;
    }
    std::vector<int> row_of(B.nd.size(), -1);
    for (int i = 0; i < n_nodes; ++i) row_of[size_t(order[i])] = i;
    memset(bits, 0, size_t(n_rows) * size_t(pitch_words) * sizeof(uint32_t));
    for (int i = 0; i < n_rows; ++i) children[2 * i] = children[2 * i + 1] = -1;
    for (int i = n_nodes - 1; i >= 0; --i) {            // leaves first, parents as the OR of their children
        const TNode& v = B.nd[size_t(order[i])];
        uint32_t* row = bits + size_t(i) * pitch_words;
        if (v.ch.empty()) {
            const int col = v.col >= 0 ? v.col : samples.find(B.text.data() + v.name_lo, size_t(v.name_len));
            if (col < 0) {
                *err = "leaf " + std::string(B.text.data() + v.name_lo, size_t(v.name_len)) + " is not a VCF sample";
                return false;
            }
            row[col >> 5] |= 1u << (col & 31);
        } else {
            if (v.ch.size() != 2) {
                *err = "the PT test needs a binary tree";
                return false;
            }
            for (int k = 0; k < 2; ++k) {
                const int cr = row_of[size_t(v.ch[size_t(k)])];
                children[2 * i + k] = cr;
                const uint32_t* crow = bits + size_t(cr) * pitch_words;
                for (int64_t w = 0; w < pitch_words; ++w) row[w] |= crow[w];
            }
        }
    }
    info[0] = n_nodes;
    info[1] = n_leaves;
    info[2] = info[3] = 0;
    return true;
}

}  // namespace
}  // namespace scs

using namespace scs;

extern "C" {

int scs_tree_parse_batch(int32_t n_trees, const char* text, const int64_t* text_off, int32_t strip_tags, int32_t n_samples,
                         const char* sample_names, const int64_t* name_off, const char* outgroup, int32_t n_rows,
                         int64_t pitch_words, uint32_t* bits, int32_t* children, int32_t* info, int32_t threads) {
    if (!text || !text_off || !sample_names || !name_off || !outgroup || !bits || !children || !info) return set_error(SCS_ERR_ARG, "null argument");
    if (n_trees <= 0 || n_samples <= 0 || n_rows <= 0 || pitch_words < (n_samples + 31) / 32)
        return set_error(SCS_ERR_ARG, "bad shape (trees=%d samples=%d rows=%d pitch=%lld)", n_trees, n_samples, n_rows, (long long)pitch_words);
    SampleIndex samples;
    samples.build(sample_names, name_off, n_samples);
    const std::string outg(outgroup);
    int nt = threads > 0 ? threads : int(std::max(1u, std::min(16u, std::thread::hardware_concurrency())));
    nt = std::max(1, std::min(nt, n_trees));
    std::atomic<int> first_bad(n_trees);
    std::vector<std::string> errs(static_cast<size_t>(nt), std::string());
    std::vector<int> bad_at(static_cast<size_t>(nt), -1);
    auto work = [&](int tid) {
        const int lo = int(int64_t(n_trees) * tid / nt), hi = int(int64_t(n_trees) * (tid + 1) / nt);
        for (int t = lo; t < hi; ++t) {
            if (t > first_bad.load(std::memory_order_relaxed)) break;
            std::string err;
            if (!process_tree(text + text_off[t], size_t(text_off[t + 1] - text_off[t]), strip_tags != 0, samples, outg, n_rows, pitch_words,
                              bits + size_t(t) * n_rows * pitch_words, children + size_t(t) * n_rows * 2, info + size_t(t) * 4, &err)) {
                errs[size_t(tid)] = err;
                bad_at[size_t(tid)] = t;
                int cur = first_bad.load();
                while (t < cur && !first_bad.compare_exchange_weak(cur, t)) {
                }
                break;
            }
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int i = 0; i < nt; ++i) pool.emplace_back(work, i);
        for (auto& th : pool) th.join();
    }
    const int fb = first_bad.load();
    if (fb < n_trees) {
        for (int i = 0; i < nt; ++i)
            if (bad_at[size_t(i)] == fb) return set_error(SCS_ERR_PARSE, "tree %d: %s", fb, errs[size_t(i)].c_str());
        return set_error(SCS_ERR_PARSE, "tree %d: malformed", fb);
    }
    return SCS_OK;
}

}  // extern "C"
