"""The Criteo-1TB preprocessing chain of the reference's ``datasets/scripts/nvt`` (TSV -> parquet -> hashed / log-scaled parquet ->
row-major binary -> per-column binary files for the NVT-layout dataloader) without NVTabular / dask: pyarrow streams the files record
batch by record batch, numpy does the transforms, a process pool fans out over files. Same command lines, same directory layout,
same file formats."""
