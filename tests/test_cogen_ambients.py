"""CPU: CogenEnv ambient data — the minimal xlsx reader, the packaged real table and the loader that built it."""
import os
import zipfile

import numpy as np
import pytest

from sustaingym_b200.data import cogen_ambients as CA
from sustaingym_b200.data.xlsx_reader import excel_serial_to_datetime, parse_us_date, read_xlsx

SHEET = """<?xml version="1.0" encoding="UTF-8"?>
<worksheet xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main"><sheetData>
<row r="1"><c r="A1" t="s"><v>0</v></c><c r="B1" t="s"><v>1</v></c><c r="D1" t="inlineStr"><is><t>inline</t></is></c></row>
<row r="3"><c r="A3"><v>44317.25</v></c><c r="B3"><v>1.5</v></c><c r="C3" t="b"><v>1</v></c></row>
</sheetData></worksheet>"""


def test_xlsx_reader_on_a_handmade_workbook(tmp_path):
    path = str(tmp_path / "w.xlsx")
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("xl/workbook.xml", '<workbook xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main" '
                   'xmlns:r="http://schemas.openxmlformats.org/officeDocument/2006/relationships"><sheets>'
                   '<sheet name="Data" sheetId="1" r:id="rId1"/></sheets></workbook>')
        z.writestr("xl/_rels/workbook.xml.rels", '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   '<Relationship Id="rId1" Target="worksheets/sheet1.xml"/></Relationships>')
        z.writestr("xl/sharedStrings.xml", '<sst xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main">'
                   '<si><t>Timestamp</t></si><si><r><t>Val</t></r><r><t>ue</t></r></si></sst>')
        z.writestr("xl/worksheets/sheet1.xml", SHEET)
    rows = read_xlsx(path)["Data"]
    assert rows[0] == ["Timestamp", "Value", None, "inline"]
    assert rows[1] == []                                       # row 2 is absent from the XML
    assert rows[2] == [44317.25, 1.5, True]
    assert str(excel_serial_to_datetime(rows[2][0])) == "2021-05-01 06:00:00"
    assert str(parse_us_date("12/9/2022")) == "2022-12-09"


def test_packaged_real_ambients():
    tab = CA.real_ambients()
    assert tab.shape == (253, 96, 7) and tab.dtype == np.float64 and np.isfinite(tab).all()
    lo, hi = tab.min(axis=(0, 1)), tab.max(axis=(0, 1))
    # inside the observation-space bounds the reference derived from this data (cogen/env.py:136-142)
    assert (lo >= [32, 14, 0, 0, 0, 0, 0]).all() and (hi <= [115, 15, 1, 700, 1300, 1500, 7]).all()
    assert (tab[..., 5].reshape(-1, 4).std(axis=1) == 0).mean() > 0.99          # hourly price, forward-filled to 15 min
    assert (np.diff(tab[..., 6], axis=1) == 0).all()           # one gas price per day
    wind = CA.real_ambients(100.0)
    assert (wind[..., 3] <= tab[..., 3]).all() and (wind[..., 3] >= 0).all() and (wind[..., 3] < tab[..., 3]).mean() > 0.5
    np.testing.assert_array_equal(np.delete(wind, 3, axis=2), np.delete(tab, 3, axis=2))


def test_loader_reproduces_the_packaged_table():
    data_dir = "/root/reference/sustaingym/data/cogen/ambients_data"
    if not os.path.isdir(data_dir):
        pytest.skip("reference data tree not present")
    from sustaingym_b200.data.load_ambients import construct_ambients
    np.testing.assert_array_equal(construct_ambients(data_dir, 25.0), CA.real_ambients(25.0))
