"""Minimal reader for .xlsx workbooks (Office Open XML) — enough for the reference's packaged CogenEnv data.

``openpyxl`` (what ``pandas.read_excel`` needs) is not part of the engine's environment, and the three
workbooks under ``sustaingym/data/cogen/ambients_data`` only hold plain cells: shared strings, inline
strings and numbers (dates are numbers: days since 1899-12-30).  A workbook is a zip of XML parts; this
reads ``xl/workbook.xml`` (sheet names), ``xl/sharedStrings.xml`` and the sheets, streaming each sheet
with ``iterparse`` (the operating-data sheet is 29 MB of XML).
"""
from __future__ import annotations

import datetime as _dt
import re
import xml.etree.ElementTree as ET
import zipfile

_NS = "{http://schemas.openxmlformats.org/spreadsheetml/2006/main}"
_REL = "{http://schemas.openxmlformats.org/officeDocument/2006/relationships}"
_EPOCH = _dt.datetime(1899, 12, 30)


def excel_serial_to_datetime(x: float) -> _dt.datetime:
    """Days since 1899-12-30 (the 1900 date system) -> naive datetime, rounded to the nearest second."""
    return _EPOCH + _dt.timedelta(seconds=round(float(x) * 86400.0))


def _col_index(ref: str) -> int:
    n = 0
    for ch in ref:
        if ch.isalpha():
            n = n * 26 + (ord(ch.upper()) - 64)
        else:
            break
    return n - 1


def read_xlsx(path: str) -> dict[str, list[list]]:
    """{sheet name: rows}; a row is a list of cell values (str, float or None), padded to its last used column."""
    with zipfile.ZipFile(path) as z:
        names = set(z.namelist())
        shared: list[str] = []
        if "xl/sharedStrings.xml" in names:
            with z.open("xl/sharedStrings.xml") as f:
                for _, el in ET.iterparse(f):
                    if el.tag == _NS + "si":
                        shared.append("".join(t.text or "" for t in el.iter(_NS + "t")))
                        el.clear()
        wb = ET.fromstring(z.read("xl/workbook.xml"))
        rels = ET.fromstring(z.read("xl/_rels/workbook.xml.rels"))
        target = {r.get("Id"): r.get("Target") for r in rels}
        out: dict[str, list[list]] = {}
        for sh in wb.iter(_NS + "sheet"):
            part = target[sh.get(_REL + "id")]
            part = part.lstrip("/") if part.startswith("/") else "xl/" + part
            rows: list[list] = []
            with z.open(part) as f:
                for _, el in ET.iterparse(f):
                    if el.tag != _NS + "row":
                        continue
                    row: list = []
                    for c in el.iter(_NS + "c"):
                        i = _col_index(c.get("r", "")) if c.get("r") else len(row)
                        t = c.get("t")
                        v = c.find(_NS + "v")
                        if t == "inlineStr":
                            val = "".join(x.text or "" for x in c.iter(_NS + "t"))
                        elif v is None or v.text is None:
                            val = None
                        elif t == "s":
                            val = shared[int(v.text)]
                        elif t in ("str", "e"):
                            val = v.text
                        elif t == "b":
                            val = bool(int(v.text))
                        else:
                            val = float(v.text)
                        if i >= len(row):
                            row.extend([None] * (i + 1 - len(row)))
                        row[i] = val
                    r_attr = el.get("r")
                    if r_attr is not None:                     # rows that are absent from the XML are empty rows
                        while len(rows) < int(r_attr) - 1:
                            rows.append([])
                    rows.append(row)
                    el.clear()
            out[sh.get("name")] = rows
        return out


def parse_us_date(s: str) -> _dt.date:
    """'MM/DD/YYYY' (ERCOT, EIA exports)."""
    m = re.fullmatch(r"\s*(\d{1,2})/(\d{1,2})/(\d{4})\s*", s)
    if not m:
        raise ValueError(f"not a MM/DD/YYYY date: {s!r}")
    return _dt.date(int(m.group(3)), int(m.group(1)), int(m.group(2)))
