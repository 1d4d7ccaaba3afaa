"""Dataset ingestion -> CSR: the host-side tool that produces what pctsea_matrix_load_csr consumes.

The reference loads the Human Cell Landscape into MongoDB (one `Expression` document per positive count, one
`SingleCell` document per cell) with pctsea-cl's `HumanSingleCellsDatasetCreation`
(pctsea-cl/src/main/java/edu/scripps/yates/pctsea/db/datasets/singlecellshuman/HumanSingleCellsDatasetCreation.java:142-262)
and the annotation reader `SingleCellsMetaInformationReader`
(pctsea-core/src/main/java/edu/scripps/yates/pctsea/utils/SingleCellsMetaInformationReader.java:107-214), and streams
those documents back for every analysis. Here the same source files are read once into a CSR matrix with the same
conventions, so that cell index = DB order (the ranking tie-break) and the scoring sees the same values:

  * expression file: gzip CSV, first line = cell names (R style: the header is shifted one column to the left of the
    data rows; an empty header cell is skipped), then one row per gene: name, then one short count per cell;
  * gene names upper-cased, quotes stripped; only counts > 0 become entries (zeros are skipped), a negative count is an
    error ("Maybe we need something else than a short");
  * cells are numbered in order of first appearance of a positive count (row-major over the file), which is the order
    their SingleCell documents are inserted, hence the order findByDatasetTag returns them;
  * a (cell, gene) pair seen twice keeps the LAST value (model/SingleCell.java:80);
  * cell type / biomaterial come from the annotation CSVs (columns cell_id, celltype, biomaterial; lower-cased, trimmed).
"""
from __future__ import annotations

import csv
import glob
import gzip
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def read_annotation_folder(folder: str) -> Dict[str, Tuple[str, str]]:
    """cell name -> (celltype, biomaterial) from every *.csv in the folder."""
    out: Dict[str, Tuple[str, str]] = {}
    for path in sorted(glob.glob(os.path.join(folder, "*.csv"))):
        with open(path, "r") as fh:
            idx = None
            for line in fh:
                sp = line.rstrip("\n").rstrip("\r").split(",")
                if idx is None:
                    idx = {h.lower().strip().replace('"', ""): i for i, h in enumerate(sp)}
                    if "cell_id" not in idx:
                        raise ValueError(f"File {path} has different header!")
                    continue
                if idx["cell_id"] >= len(sp):
                    continue
                name = sp[idx["cell_id"]].replace('"', "")
                ctype = sp[idx["celltype"]].strip().lower().replace('"', "") if "celltype" in idx else ""
                bio = sp[idx["biomaterial"]].strip().lower().replace('"', "") if "biomaterial" in idx else ""
                out[name] = (ctype, bio)
    return out


class CSRDataset:
    def __init__(self, row_ptr, col, val, gene_names, cell_names, cell_types):
        self.row_ptr, self.col, self.val = row_ptr, col, val
        self.gene_names, self.cell_names, self.cell_types = gene_names, cell_names, cell_types

    @property
    def shape(self):
        return len(self.cell_names), len(self.gene_names)

    def save(self, path: str):
        np.savez_compressed(path, row_ptr=self.row_ptr, col=self.col, val=self.val,
                            gene_names=np.asarray(self.gene_names, dtype=object),
                            cell_names=np.asarray(self.cell_names, dtype=object),
                            cell_types=np.asarray(self.cell_types, dtype=object))

    @classmethod
    def load(cls, path: str) -> "CSRDataset":
        z = np.load(path, allow_pickle=True)
        return cls(z["row_ptr"], z["col"], z["val"], list(z["gene_names"]), list(z["cell_names"]), list(z["cell_types"]))

    # ---- the binary dump shared with the Java side (java/src/.../gpu/MongoToCsr.java writes it from the reference's
    # MongoDB, GpuDataset.java reads it): "<base>.csr" little endian = magic "PCTSEACSR1\0\0" | nCells i64 | nGenes i32 |
    # nTypes i32 | nnz i64 | rowPtr i64[nCells+1] | col i32[nnz] | val f32[nnz] | label i32[nCells], plus
    # "<base>.genes.txt" / ".types.txt" / ".cells.txt" with one name per line (line i = column / label id / row i).
    MAGIC = b"PCTSEACSR1\0\0"

    def labels(self):
        """(label int32[nCells], type names): ids in order of first appearance, -1 for a cell without a type."""
        ids: Dict[str, int] = {}
        lab = np.empty(len(self.cell_types), np.int32)
        for i, t in enumerate(self.cell_types):
            if t is None or t == "":
                lab[i] = -1
            else:
                lab[i] = ids.setdefault(t, len(ids))
        return lab, list(ids.keys())

    def save_binary(self, base: str):
        lab, types = self.labels()
        n_cells, n_genes = self.shape
        with open(base + ".csr", "wb") as fh:
            fh.write(self.MAGIC)
            fh.write(np.array([n_cells], "<i8").tobytes())
            fh.write(np.array([n_genes, len(types)], "<i4").tobytes())
            fh.write(np.array([len(self.col)], "<i8").tobytes())
            fh.write(np.ascontiguousarray(self.row_ptr, "<i8").tobytes())
            fh.write(np.ascontiguousarray(self.col, "<i4").tobytes())
            fh.write(np.ascontiguousarray(self.val, "<f4").tobytes())
            fh.write(np.ascontiguousarray(lab, "<i4").tobytes())
        for suffix, names in ((".genes.txt", self.gene_names), (".types.txt", types), (".cells.txt", self.cell_names)):
            with open(base + suffix, "w") as fh:
                fh.writelines(str(n) + "\n" for n in names)

    @classmethod
    def load_binary(cls, base: str) -> "CSRDataset":
        with open(base + ".csr", "rb") as fh:
            if fh.read(12) != cls.MAGIC:
                raise ValueError(f"{base}.csr is not a PCTSEA CSR dump")
            n_cells = int(np.frombuffer(fh.read(8), "<i8")[0])
            n_genes, n_types = (int(v) for v in np.frombuffer(fh.read(8), "<i4"))
            nnz = int(np.frombuffer(fh.read(8), "<i8")[0])
            row_ptr = np.frombuffer(fh.read(8 * (n_cells + 1)), "<i8").astype(np.int64)
            col = np.frombuffer(fh.read(4 * nnz), "<i4").astype(np.int32)
            val = np.frombuffer(fh.read(4 * nnz), "<f4").astype(np.float32)
            lab = np.frombuffer(fh.read(4 * n_cells), "<i4")
        read = lambda suffix: [ln.rstrip("\n") for ln in open(base + suffix)]
        genes, types, cells = read(".genes.txt"), read(".types.txt"), read(".cells.txt")
        if len(genes) != n_genes or len(types) != n_types or len(cells) != n_cells or int(row_ptr[-1]) != nnz:
            raise ValueError(f"{base}: name files do not match the CSR header")
        return cls(row_ptr, col, val, genes, cells, [types[t] if t >= 0 else "" for t in lab])


def read_hcl_expression_files(files: Sequence[str], annotations: Optional[Dict[str, Tuple[str, str]]] = None) -> CSRDataset:
    cell_index: Dict[str, int] = {}
    gene_index: Dict[str, int] = {}
    cell_names: List[str] = []
    gene_names: List[str] = []
    entries: Dict[Tuple[int, int], float] = {}      # (cell, gene) -> value; re-assignment = last document wins
    for path in files:
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rt") as fh:
            headers = None
            for line in fh:
                line = line.rstrip("\n").rstrip("\r")
                if line.startswith("#"):
                    continue
                sp = line.split(",")
                if headers is None:
                    headers = [h.replace('"', "") for h in sp]   # header i <-> data column i + 1
                    continue
                gene = sp[0].replace('"', "").upper()
                row_cells, row_vals = [], []
                for i, header in enumerate(headers):
                    if header == "":
                        continue
                    v = int(sp[i + 1])
                    if not -32768 <= v <= 32767:
                        raise ValueError(f"Value out of range for a short: {sp[i + 1]}")
                    if v > 0:
                        c = cell_index.get(header)
                        if c is None:
                            c = len(cell_names)
                            cell_index[header] = c
                            cell_names.append(header)
                        row_cells.append(c)
                        row_vals.append(float(v))
                    elif v < 0:
                        raise ValueError("Maybe we need something else than a short for this: " + sp[i + 1])
                if row_cells:
                    g = gene_index.get(gene)
                    if g is None:
                        g = len(gene_names)
                        gene_index[gene] = g
                        gene_names.append(gene)
                    for c, v in zip(row_cells, row_vals):
                        entries[(c, g)] = v
    n_cells = len(cell_names)
    keys = np.array(list(entries.keys()), dtype=np.int64).reshape(-1, 2)
    vals = np.array(list(entries.values()), dtype=np.float32)
    order = np.lexsort((keys[:, 1], keys[:, 0])) if len(keys) else np.zeros(0, np.int64)
    keys, vals = keys[order], vals[order]
    row_ptr = np.zeros(n_cells + 1, np.int64)
    if len(keys):
        np.add.at(row_ptr, keys[:, 0] + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    types = [(annotations.get(n, ("", ""))[0] if annotations else "") for n in cell_names]
    return CSRDataset(row_ptr, keys[:, 1].astype(np.int32) if len(keys) else np.zeros(0, np.int32), vals,
                      gene_names, cell_names, types)


def ingest_hcl(expression_folder: str, annotation_folder: Optional[str] = None) -> CSRDataset:
    """PCTSEADbApplication `HCL <exprDir> <annotDir>` (PCTSEADbApplication.java:171-199), producing a CSR instead of
    MongoDB collections."""
    ann = read_annotation_folder(annotation_folder) if annotation_folder else None
    files = sorted(glob.glob(os.path.join(expression_folder, "*.gz")))
    return read_hcl_expression_files(files, ann)
