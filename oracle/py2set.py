"""CPython-2.7 ``list(set_of_str)`` iteration order (TEST INFRASTRUCTURE ONLY).

generate_peak_matrix/__init__.py:44 builds the barcode universe as
``list({bc for ... in open_fragment_file(...)})``; the matrix column order is
therefore the slot order of a Python-2.7 set of ``str`` filled in order of first
appearance.  Restated from CPython 2.7 ``Objects/stringobject.c:string_hash`` and
``Objects/setobject.c`` (``set_lookkey_string``, ``set_insert_clean``,
``set_add_key``, ``set_table_resize``; PySet_MINSIZE 8, PERTURB_SHIFT 5), hash
randomisation off (the 2.7 default).  No python2 exists in the build container,
so this order is "parity unpinned" (level L2); the
required level L1 (columns matched by barcode string) does not depend on it.
"""
MASK64 = (1 << 64) - 1


def py2_str_hash(s):
    """string_hash: 64-bit signed wrap-around, -1 mapped to -2, '' hashes to 0."""
    if isinstance(s, str):
        s = s.encode("latin-1")
    if len(s) == 0:
        return 0
    x = (s[0] << 7) & MASK64
    for c in s:
        x = ((1000003 * x) & MASK64) ^ c
    x ^= len(s)
    if x >= 1 << 63:
        x -= 1 << 64
    if x == -1:
        x = -2
    return x


class Py2Set(object):
    def __init__(self):
        self.mask = 7
        self.table = [None] * 8      # entries: (key, hash)
        self.used = 0
        self.fill = 0

    def _probe(self, key, h, clean):
        mask = self.mask
        i = h & mask                  # python ints: & with a non-negative mask acts on two's complement
        perturb = h & MASK64
        e = self.table[i]
        if e is None or (not clean and e[1] == h and e[0] == key):
            return i
        while True:
            i = (i * 5 + perturb + 1) & MASK64
            e = self.table[i & mask]
            if e is None or (not clean and e[1] == h and e[0] == key):
                return i & mask
            perturb >>= 5

    def add(self, key):
        h = py2_str_hash(key)
        slot = self._probe(key, h, clean=False)
        if self.table[slot] is not None:
            return
        self.table[slot] = (key, h)
        self.used += 1
        self.fill += 1
        if self.fill * 3 >= (self.mask + 1) * 2:
            self._resize(self.used * 2 if self.used > 50000 else self.used * 4)

    def _resize(self, minused):
        newsize = 8
        while newsize <= minused:
            newsize <<= 1
        old = self.table
        self.table = [None] * newsize
        self.mask = newsize - 1
        for e in old:
            if e is not None:
                slot = self._probe(e[0], e[1], clean=True)
                self.table[slot] = e
        self.fill = self.used

    def ordered(self):
        return [e[0] for e in self.table if e is not None]


def py2_set_order(keys_in_first_appearance_order):
    s = Py2Set()
    for k in keys_in_first_appearance_order:
        s.add(k)
    return s.ordered()
