"""A minimal in-process stand-in for the `lmdb` module (absent in this image), just enough of its API for
LOPQSearcherLMDB (search.py:385-499 in the reference): lmdb.open(path, map_size=, max_dbs=) -> env; env.open_db(name);
env.begin(db=, write=) -> transaction usable as a context manager with put / get / cursor() (iteration in KEY ORDER, as
LMDB's B-tree gives it) / commit; env.sync().  "Databases" live in a module-level dict keyed by path, so that a second
environment opened on the same path sees what the first one wrote -- the persistence the product relies on."""
_STORES = {}


class _Txn(object):
    def __init__(self, store, write):
        self.store, self.write = store, write

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def put(self, key, value):
        assert self.write
        self.store[bytes(key)] = bytes(value)
        return True

    def get(self, key, default=None):
        return self.store.get(bytes(key), default)

    def cursor(self):
        return iter(sorted(self.store.items()))

    def commit(self):
        pass

    def stat(self):
        return {"entries": len(self.store)}


class _Env(object):
    def __init__(self, path):
        self.path = path
        self.dbs = _STORES.setdefault(path, {})

    def open_db(self, name):
        self.dbs.setdefault(bytes(name), {})
        return bytes(name)

    def begin(self, db=None, write=False):
        return _Txn(self.dbs[db], write)

    def sync(self):
        pass

    def close(self):
        pass


def open(path, map_size=0, max_dbs=0):      # noqa: A001  (the module's public name)
    return _Env(path)
