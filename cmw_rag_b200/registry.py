"""One HBM-resident collection per product version.

Mirrors how the reference picks and caches its stores: ``get_collection_name(version)``
(rag_engine/config/settings.py:261-273: the per-version override if non-empty, else ``{base}_v5`` /
``{base}_v6``, the base name unchanged for an unknown version) and the per-version retriever cache of
``rag_engine/tools/retrieve_context.py:101-131`` (one ``ChromaStore(collection_name=...)`` per version, created on
first use, kept for the life of the process).  Where the Chroma server finds a collection in its ``--path``
directory (systemd/cmw-rag-chroma.service:11), the registry loads ``<root>/<collection name>/`` written by
``B200Store.save``.
"""
from __future__ import annotations

import os
import threading
from typing import Callable


class CollectionRegistry:
    def __init__(self, base_collection: str = "default", overrides: dict[str, str] | None = None,
                 versions: tuple[str, ...] = ("v5", "v6"), root: str | None = None,
                 factory: Callable | None = None, **store_kwargs):
        """``overrides``: version -> collection name (the reference's ``chromadb_collection_v5`` / ``_v6``
        settings; empty values fall back to ``{base}_{version}``).  ``root``: directory of saved collections.
        ``factory(collection_name, **store_kwargs)`` builds an empty store (default: ``B200Store``)."""
        self.base = base_collection
        self.overrides = dict(overrides or {})
        self.versions = tuple(versions)
        self.root = root
        self._factory = factory
        self._kw = store_kwargs
        self._stores: dict[str, object] = {}
        self._lock = threading.Lock()

    def collection_name(self, version: str | None) -> str:
        """settings.py:261-273, for any configured version instead of the two hard-coded ones."""
        if version in self.versions:
            return self.overrides.get(version) or f"{self.base}_{version}"
        return self.base

    def path_of(self, version: str | None) -> str | None:
        return os.path.join(self.root, self.collection_name(version)) if self.root else None

    def get_store(self, version: str | None = None):
        """The store of a product version: cached; else loaded from ``<root>/<name>/`` if saved there; else empty."""
        name = self.collection_name(version)
        with self._lock:
            st = self._stores.get(name)
            if st is None:
                st = self._open(name)
                self._stores[name] = st
            return st

    def _open(self, name: str):
        from .store import B200Store

        path = os.path.join(self.root, name) if self.root else None
        if path and os.path.exists(os.path.join(path, "meta.json")):
            kw = {k: v for k, v in self._kw.items() if k in ("capacity", "device", "keep_f32", "keep_bf16", "tiles16")}
            return B200Store.load(path, **kw)
        if self._factory is not None:
            return self._factory(name, **self._kw)
        return B200Store(collection_name=name, **self._kw)

    def save(self, version: str | None = None) -> str:
        if not self.root:
            raise ValueError("CollectionRegistry has no root directory")
        path = self.path_of(version)
        self.get_store(version).save(path)
        return path

    def loaded(self) -> dict[str, object]:
        with self._lock:
            return dict(self._stores)

    def close(self) -> None:
        with self._lock:
            for st in self._stores.values():
                close = getattr(st, "close", None)
                if close:
                    close()
            self._stores.clear()
