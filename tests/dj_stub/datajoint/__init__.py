"""In-memory stand-in for the DataJoint API subset the reference uses (SURVEY §4 "Fake DataJoint", §5 config row), so that
the reference's REAL ``pose_pipeline/pipeline.py`` and ``utils/standard_pipelines.py`` import and run without MySQL:

    dj.config, dj.schema, dj.Manual / Lookup / Computed / Imported / Part, ``definition`` parsing (primary key above ``---``,
    ``-> Parent`` references, defaults), ``contents`` of Lookup tables, insert1 / insert (skip_duplicates,
    allow_direct_insert, referential integrity), fetch / fetch1 ("KEY", attribute lists, as_dict), restriction ``&`` by
    dict / SQL-ish string / another relation / list, antijoin ``-``, natural join ``*``, proj, len, delete, key_source,
    populate(*restrictions, reserve_jobs, suppress_errors), ``attach@store`` attributes (fetch returns a fresh local copy).

Test infrastructure only; nothing under posepipeline_b200/ imports it.
"""
from __future__ import annotations

import copy
import os
import re
import shutil
import tempfile

import numpy as np

__version__ = "0.14.stub"


class DataJointError(Exception):
    pass


class DuplicateError(DataJointError):
    pass


class IntegrityError(DataJointError):
    pass


class _Config(dict):
    def save_local(self, *a, **k):
        pass

    def save_global(self, *a, **k):
        pass


config = _Config({"database.host": "stub", "enable_python_native_blobs": True, "stores": {}, "custom": {}})
_registry = {}                       # class name -> table class (all schemas)
_attach_dir = tempfile.mkdtemp(prefix="djstub_store_")
_download_dir = tempfile.mkdtemp(prefix="djstub_dl_")


def conn(*a, **k):
    return None


# ------------------------------------------------------------------------------------------ restrictions
_COND = re.compile(r"""^\s*(\w+)\s*(>=|<=|!=|<>|=|>|<)\s*("[^"]*"|'[^']*'|[-+.\w]+)\s*$""")


def _parse_condition(s):
    terms = []
    for part in re.split(r"\s+AND\s+", s, flags=re.I):
        m = _COND.match(part)
        if not m:
            raise DataJointError(f"dj stub: unsupported restriction string {s!r}")
        name, op, val = m.groups()
        if val[0] in "\"'":
            val = val[1:-1]
        else:
            try:
                val = int(val)
            except ValueError:
                val = float(val)
        terms.append((name, op, val))
    return terms


def _cmp(a, op, b):
    if isinstance(b, (int, float)) and not isinstance(a, (int, float, np.integer, np.floating)):
        return False
    return {"=": a == b, "!=": a != b, "<>": a != b, ">": a > b, "<": a < b, ">=": a >= b, "<=": a <= b}[op]


def _eq(a, b):
    try:
        return bool(a == b)
    except Exception:
        return False


class QueryExpression:
    """A set of rows (list of dicts) with a heading (primary key + secondary attribute names)."""

    def __init__(self, primary_key, secondary, rows, table=None):
        self.primary_key, self.secondary, self._rows, self._table = list(primary_key), list(secondary), rows, table

    # -- heading
    @property
    def heading_names(self):
        return self.primary_key + self.secondary

    def _rows_now(self):
        return self._rows

    # -- operators
    def _restrict_rows(self, r):
        rows = self._rows_now()
        if r is None or (isinstance(r, bool) and r):
            return rows
        if isinstance(r, dict):
            keys = [k for k in r if k in self.heading_names]
            return [row for row in rows if all(_eq(row.get(k), r[k]) for k in keys)]
        if isinstance(r, str):
            terms = _parse_condition(r)
            for name, _, _ in terms:
                if name not in self.heading_names:
                    raise DataJointError(f"Unknown column '{name}' in restriction")
            return [row for row in rows if all(_cmp(row.get(n), op, v) for n, op, v in terms)]
        if isinstance(r, (list, tuple)):
            out = []
            for row in rows:
                if any(row in QueryExpression(self.primary_key, self.secondary, [row])._restrict_rows(x) for x in r):
                    out.append(row)
            return out
        other = _as_query(r)
        if other is not None:
            common = [k for k in self.heading_names if k in other.heading_names]
            orows = other._rows_now()
            return [row for row in rows if any(all(_eq(row.get(k), o.get(k)) for k in common) for o in orows)]
        raise DataJointError(f"dj stub: unsupported restriction {type(r)}")

    def __and__(self, r):
        return QueryExpression(self.primary_key, self.secondary, self._restrict_rows(r), self._table)

    def __sub__(self, r):
        other = _as_query(r)
        if other is None:
            keep = self._restrict_rows(r)
            rows = [row for row in self._rows_now() if not any(row is k for k in keep)]
        else:
            common = [k for k in self.heading_names if k in other.heading_names]
            orows = other._rows_now()
            rows = [row for row in self._rows_now() if not any(all(_eq(row.get(k), o.get(k)) for k in common) for o in orows)]
        return QueryExpression(self.primary_key, self.secondary, rows, self._table)

    def __mul__(self, r):
        other = _as_query(r)
        common = [k for k in self.heading_names if k in other.heading_names]
        rows = []
        for a in self._rows_now():
            for b in other._rows_now():
                if all(_eq(a.get(k), b.get(k)) for k in common):
                    rows.append({**b, **a})
        pk = self.primary_key + [k for k in other.primary_key if k not in self.primary_key]
        sec = [k for k in self.secondary + other.secondary if k not in pk]
        return QueryExpression(pk, list(dict.fromkeys(sec)), rows)

    def proj(self, *attrs, **renamed):
        if renamed:
            raise DataJointError("dj stub: renaming projection not supported")
        keep = self.primary_key + [a for a in attrs if a in self.secondary]
        rows = [{k: row[k] for k in keep} for row in self._rows_now()]
        return QueryExpression(self.primary_key, [a for a in attrs if a in self.secondary], rows)

    def __len__(self):
        return len(self._rows_now())

    def __bool__(self):
        return len(self) > 0

    def __iter__(self):
        return iter(self.fetch(as_dict=True))

    # -- fetch
    def _value(self, row, attr):
        v = row[attr]
        if isinstance(v, _Attachment):
            return v.download()
        return copy.deepcopy(v)

    def fetch(self, *attrs, as_dict=False, order_by=None, limit=None, download_path=None, squeeze=False):
        rows = list(self._rows_now())
        if order_by:
            keys = [order_by] if isinstance(order_by, str) else list(order_by)
            for k in reversed(keys):
                name, _, direction = k.partition(" ")
                rows.sort(key=lambda r: r[name] if name != "KEY" else tuple(r[p] for p in self.primary_key),
                          reverse=direction.strip().upper() == "DESC")
        if limit is not None:
            rows = rows[:limit]
        if not attrs:
            out = [{k: self._value(r, k) for k in self.heading_names} for r in rows]
            return out
        cols = []
        for a in attrs:
            if a == "KEY":
                cols.append([{k: r[k] for k in self.primary_key} for r in rows])
            else:
                if a not in self.heading_names:
                    raise DataJointError(f"Attribute `{a}` not found")
                cols.append([self._value(r, a) for r in rows])
        if as_dict:
            out = []
            for i in range(len(rows)):
                d = {}
                for a, c in zip(attrs, cols):
                    if a == "KEY":
                        d.update(c[i])
                    else:
                        d[a] = c[i]
                out.append(d)
            return out

        def arr(a, c):
            if a == "KEY":
                return c
            o = np.empty(len(c), dtype=object)
            for i, v in enumerate(c):
                o[i] = v
            try:
                if all(np.isscalar(v) for v in c) and c:
                    return np.array(c)
            except Exception:
                pass
            return o
        res = [arr(a, c) for a, c in zip(attrs, cols)]
        return res[0] if len(res) == 1 else res

    def fetch1(self, *attrs, **kw):
        rows = self._rows_now()
        if len(rows) != 1:
            raise DataJointError(f"fetch1 should only return one tuple. {len(rows)} tuples found")
        r = rows[0]
        if not attrs:
            return {k: self._value(r, k) for k in self.heading_names}
        vals = tuple({k: r[k] for k in self.primary_key} if a == "KEY" else self._value(r, a) for a in attrs)
        for a in attrs:
            if a != "KEY" and a not in self.heading_names:
                raise DataJointError(f"Attribute `{a}` not found")
        return vals[0] if len(vals) == 1 else vals

    def delete(self, *a, **k):
        if self._table is None:
            raise DataJointError("cannot delete from a derived relation")
        doomed = self._rows_now()
        self._table._store[:] = [r for r in self._table._store if not any(r is d for d in doomed)]

    delete_quick = delete


def _as_query(x):
    if isinstance(x, QueryExpression):
        return x
    if isinstance(x, type) and issubclass(x, Table):
        return x()
    return None


class _Attachment:
    """attach@store attribute: the file content lives in the stub's store; every fetch materialises a fresh local copy
    (DataJoint downloads the attachment; the reference then ``shutil.move``s it, pipeline.py:53-56)."""

    def __init__(self, path):
        self.name = os.path.basename(path)
        fd, self.stored = tempfile.mkstemp(dir=_attach_dir, suffix="_" + self.name)
        os.close(fd)
        shutil.copy(path, self.stored)

    def download(self):
        d = tempfile.mkdtemp(dir=_download_dir)
        out = os.path.join(d, self.name)
        shutil.copy(self.stored, out)
        return out


# ------------------------------------------------------------------------------------------ tables
class _TableMeta(type):
    """Like DataJoint's TableMeta: relational operators and table methods work on the class itself."""

    def __and__(cls, r):
        return cls() & r

    def __sub__(cls, r):
        return cls() - r

    def __mul__(cls, r):
        return cls() * r

    def __len__(cls):
        return len(cls())

    def __iter__(cls):
        return iter(cls())

    def __bool__(cls):
        return True

    _FORWARD = frozenset(("populate", "insert1", "insert", "fetch", "fetch1", "delete", "delete_quick", "proj", "primary_key",
                          "heading_names", "heading", "key_source", "drop", "drop_quick"))

    def __getattribute__(cls, name):
        # table methods called on the class act on an instance (DataJoint's TableMeta does the same)
        if name in _TableMeta._FORWARD and type.__getattribute__(cls, "_declared"):
            return getattr(cls(), name)
        return type.__getattribute__(cls, name)


_ATTR = re.compile(r"^(\w+)\s*(?:=\s*(.+?))?\s*:\s*([^#]+?)\s*(?:#.*)?$")


class Table(QueryExpression, metaclass=_TableMeta):
    definition = ""
    _declared = False
    _computed = False

    def __init__(self):
        cls = type(self)
        if not cls._declared:
            raise DataJointError(f"{cls.__name__} is not decorated with a schema")
        QueryExpression.__init__(self, cls._pk, cls._sec, cls._store, cls)
        self._allow_insert = not cls._computed

    def _rows_now(self):
        return type(self)._store

    @classmethod
    def _declare(cls):
        pk, sec, defaults, parents, types = [], [], {}, [], {}
        in_pk = True
        for line in cls.definition.strip().splitlines():
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            if line.startswith("---") or line.startswith("___"):
                in_pk = False
                continue
            if line.startswith("->"):
                name = re.sub(r"\[.*?\]", "", line[2:]).split("#")[0].strip()
                if name not in _registry:
                    raise DataJointError(f"{cls.__name__}: unknown parent table {name}")
                parent = _registry[name]
                parents.append((parent, in_pk))
                for a in parent._pk:
                    if a not in pk and a not in sec:
                        (pk if in_pk else sec).append(a)
                continue
            m = _ATTR.match(line)
            if not m:
                raise DataJointError(f"{cls.__name__}: cannot parse definition line {line!r}")
            name, default, typ = m.groups()
            (pk if in_pk else sec).append(name)
            types[name] = typ.strip()
            if default is not None:
                defaults[name] = default.strip()
        cls._pk, cls._sec, cls._defaults, cls._parents, cls._types = pk, sec, defaults, parents, types
        cls._store = []
        cls._declared = True

    @property
    def heading(self):
        return type("Heading", (), {"names": self.heading_names, "primary_key": self.primary_key,
                                    "secondary_attributes": self.secondary})()

    # -- insert
    def insert1(self, row, **kw):
        self.insert([row], **kw)

    def insert(self, rows, replace=False, skip_duplicates=False, ignore_extra_fields=False, allow_direct_insert=None, **kw):
        cls = type(self)
        if cls._computed and not (allow_direct_insert or getattr(cls, "_in_make", 0) > 0):
            raise DataJointError("Inserts into an auto-populated table can only be done inside its make method "
                                 "during a populate call. To override, set keyword argument allow_direct_insert=True.")
        for row in rows:
            if isinstance(row, np.void):
                row = {k: row[k] for k in row.dtype.names}
            elif not isinstance(row, dict):
                row = dict(zip(cls._pk + cls._sec, row))
            extra = [k for k in row if k not in cls._pk + cls._sec]
            if extra and not ignore_extra_fields:
                raise DataJointError(f"Field '{extra[0]}' not in the table heading of {cls.__name__}")
            rec = {}
            for a in cls._pk + cls._sec:
                if a in row and row[a] is not None:
                    v = row[a]
                    if cls._types.get(a, "").startswith("attach"):
                        v = _Attachment(v)
                    else:
                        v = copy.deepcopy(v)
                        if isinstance(v, np.generic) and np.isscalar(v):
                            v = v.item()
                    rec[a] = v
                elif a in cls._defaults:
                    d = cls._defaults[a]
                    rec[a] = None if d.lower() == "null" else d.strip("\"'")
                else:
                    raise DataJointError(f"Field '{a}' doesn't have a default value")
            dup = [r for r in cls._store if all(_eq(r[k], rec[k]) for k in cls._pk)]
            if dup:
                if skip_duplicates:
                    continue
                if replace:
                    cls._store[:] = [r for r in cls._store if r is not dup[0]]
                else:
                    raise DuplicateError(f"Duplicate entry for key PRIMARY in {cls.__name__}")
            for parent, _ in cls._parents:
                if not any(all(_eq(p[k], rec[k]) for k in parent._pk) for p in parent._store):
                    raise IntegrityError(f"Cannot add or update a child row: a foreign key constraint fails "
                                         f"({cls.__name__} -> {parent.__name__})")
            cls._store.append(rec)

    def drop(self, *a, **k):
        type(self)._store[:] = []

    drop_quick = drop


class Manual(Table):
    pass


class Lookup(Table):
    contents = []


class Part(Table):
    pass


class _AutoPopulate(Table):
    _computed = True

    @property
    def key_source(self):
        src = None
        for parent, in_pk in type(self)._parents:
            if in_pk:
                q = parent().proj()
                src = q if src is None else src * q
        if src is None:
            raise DataJointError("A relation must have primary dependencies for auto-populate to work")
        return src

    def populate(self, *restrictions, suppress_errors=False, return_exception_objects=False, reserve_jobs=False,
                 order="original", limit=None, max_calls=None, display_progress=False, processes=1, make_kwargs=None):
        cls = type(self)
        todo = _as_query(self.key_source)
        for r in restrictions:
            todo = todo & r
        todo = todo.proj() - self
        keys = todo.fetch("KEY")
        if limit is not None:
            keys = keys[:limit]
        errors = []
        for i, key in enumerate(keys):
            if max_calls is not None and i >= max_calls:
                break
            if len(self & key):
                continue
            cls._in_make = getattr(cls, "_in_make", 0) + 1
            before = len(cls._store)
            try:
                self.make(dict(key), **(make_kwargs or {}))
            except (KeyboardInterrupt, SystemExit):
                raise
            except Exception as e:
                del cls._store[before:]                    # the make() runs in a transaction: roll back partial inserts
                if not suppress_errors:
                    raise
                errors.append(e if return_exception_objects else (key, str(e)))
            finally:
                cls._in_make -= 1
        if suppress_errors:
            return errors


class Computed(_AutoPopulate):
    pass


class Imported(_AutoPopulate):
    pass


# ------------------------------------------------------------------------------------------ schema
class Schema:
    def __init__(self, schema_name=None, context=None, **kw):
        self.database, self.context, self.tables = schema_name, context, {}

    def __call__(self, cls, *, context=None):
        cls._declare()
        _registry[cls.__name__] = cls
        self.tables[cls.__name__] = cls
        if issubclass(cls, Lookup) and cls.contents:
            cls().insert(cls.contents, skip_duplicates=True)
        return cls

    def spawn_missing_classes(self, context=None):
        pass

    def drop(self, force=False):
        for t in self.tables.values():
            t._store[:] = []

    @property
    def jobs(self):
        return []


schema = Schema


def reset():
    """Empty every table (Lookup contents are re-inserted) -- test isolation."""
    for cls in _registry.values():
        cls._store[:] = []
    for cls in _registry.values():
        if issubclass(cls, Lookup) and cls.contents:
            cls().insert(cls.contents, skip_duplicates=True)
