"""MovieLens 20M / 25M rating datapipes (reference torchrec/datasets/movielens.py:37-140)."""
from __future__ import annotations

import csv
import os
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Union

from .utils import LoadFiles, ReadLinesFromCSV, safe_cast

RATINGS_FILENAME = "ratings.csv"
MOVIES_FILENAME = "movies.csv"
DEFAULT_RATINGS_COLUMN_NAMES: List[str] = ["userId", "movieId", "rating", "timestamp"]
DEFAULT_MOVIES_COLUMN_NAMES: List[str] = ["movieId", "title", "genres"]
DEFAULT_COLUMN_NAMES: List[str] = DEFAULT_RATINGS_COLUMN_NAMES + DEFAULT_MOVIES_COLUMN_NAMES[1:]
COLUMN_TYPE_CASTERS: List[Callable[[Union[float, int, str]], Union[float, int, str]]] = [
    lambda v: safe_cast(v, int, 0), lambda v: safe_cast(v, int, 0), lambda v: safe_cast(v, float, 0.0), lambda v: safe_cast(v, int, 0),
    lambda v: safe_cast(v, str, ""), lambda v: safe_cast(v, str, "")]


def _default_row_mapper(example: List[str]) -> Dict[str, Union[float, int, str]]:
    return {DEFAULT_COLUMN_NAMES[idx]: COLUMN_TYPE_CASTERS[idx](val) for idx, val in enumerate(example)}


class _MovieLens:
    def __init__(self, root: str, include_movies_data: bool, row_mapper: Optional[Callable[[List[str]], Any]], **open_kw: Any) -> None:
        self.root, self.include_movies_data, self.row_mapper, self.open_kw = root, include_movies_data, row_mapper, open_kw

    def __iter__(self) -> Iterator[Any]:
        movies: Dict[str, List[str]] = {}
        if self.include_movies_data:
            with open(os.path.join(self.root, MOVIES_FILENAME), "r", newline="", encoding="utf-8") as f:
                rd = csv.reader(f)
                next(rd, None)
                for row in rd:
                    movies[row[0]] = row[1:]
        rows = ReadLinesFromCSV(LoadFiles((os.path.join(self.root, RATINGS_FILENAME),), mode="r", **self.open_kw), skip_first_line=True, delimiter=",")
        for row in rows:
            if self.include_movies_data:
                row = list(row) + movies.get(row[1], ["", ""])
            yield self.row_mapper(row) if self.row_mapper else row


def movielens_20m(root: str, *, include_movies_data: bool = False, row_mapper: Optional[Callable[[List[str]], Any]] = _default_row_mapper, **open_kw: Any) -> Iterable:
    return _MovieLens(root, include_movies_data, row_mapper, **open_kw)


def movielens_25m(root: str, *, include_movies_data: bool = False, row_mapper: Optional[Callable[[List[str]], Any]] = _default_row_mapper, **open_kw: Any) -> Iterable:
    return _MovieLens(root, include_movies_data, row_mapper, **open_kw)
