"""Reference metadata: the subset of lib/python/tools/ref_manager.py (ReferenceManager) the three stages use.

Reads `<ref>/fasta/genome.fa.fai` (contig order and lengths, ref_manager.py:45-52), `<ref>/fasta/contig-defs.json`
(primary contigs, species prefixes, sex chromosomes, :25-43,140-159) and `<ref>/genome` (:68-73)."""
import json
import os


class ReferenceManager(object):
    def __init__(self, path):
        self.path = path
        defs = os.path.join(path, "fasta", "contig-defs.json")
        if not os.path.exists(defs):
            raise IOError("Reference is missing contig definition file")
        try:
            with open(defs, "r") as f:
                self.contigs = json.load(f)
        except ValueError as e:
            raise Exception("Malformed json in '%s': %s" % (defs, e))
        if self.contigs.get("non_nuclear_contigs", []) is None:
            self.contigs["non_nuclear_contigs"] = []
        self.sex_chromosomes = {}
        if "sex_chromosomes" in self.contigs:
            if not isinstance(self.contigs["sex_chromosomes"], dict):
                raise Exception("Malformed entry for sex_chromosomes in {}".format(defs))
            for species in self.contigs["sex_chromosomes"]:
                for contig in self.contigs["sex_chromosomes"][species]:
                    self.sex_chromosomes[contig] = True
        fai = os.path.join(path, "fasta", "genome.fa.fai")
        if not os.path.exists(fai):
            raise IOError("Reference is missing fasta index file")
        self.all_contigs = []
        self.contig_lengths = {}
        with open(fai, "r") as f:
            for line in f:
                fields = line.strip().split()
                self.all_contigs.append(fields[0])
                self.contig_lengths[fields[0]] = int(fields[1])

    @property
    def genome(self):
        p = os.path.join(self.path, "genome")
        if not os.path.exists(p):
            raise IOError("Reference is missing genome name file")
        with open(p) as f:
            return f.readline().strip()

    def get_contig_lengths(self):
        return self.contig_lengths

    def list_species(self):
        return self.contigs["species_prefixes"]

    def primary_contigs(self, species=None, allow_sex_chromosomes=False):
        return [x for x in self.contigs["primary_contigs"]
                if allow_sex_chromosomes or x not in self.sex_chromosomes]

    def contig_table(self):
        """[(name, length)] in genome.fa.fai order: the order `bedtools sort -g` gives the final peaks.bed."""
        return [(c, self.contig_lengths[c]) for c in self.all_contigs]


def generate_genome_tag(ref_path):
    """lib/python/utils.py:87-94."""
    mgr = ReferenceManager(ref_path)
    genomes = mgr.list_species()
    if (len(genomes) == 1 and genomes[0] == "") or len(genomes) == 0:
        genomes = [mgr.genome]
    return genomes


def genome_of_contig(name, genomes):
    """cellranger/atac/feature_ref.py:30-43."""
    if len(genomes) == 1:
        return genomes[0]
    for g in genomes:
        if name.startswith(g):
            return g
    raise Exception("%s does not have valid associated genome" % name)


def write_reference(path, contigs, genome="GRCh38", sex=("chrX", "chrY")):
    """Minimal reference directory for tests / synthetic runs (format of lib/python/reference.py:400-439)."""
    os.makedirs(os.path.join(path, "fasta"), exist_ok=True)
    with open(os.path.join(path, "fasta", "genome.fa.fai"), "w") as f:
        off = 6
        for name, length in contigs:
            f.write("%s\t%d\t%d\t60\t61\n" % (name, length, off))
            off += length + length // 60 + len(name) + 2
    names = [c[0] for c in contigs]
    sx = [s for s in sex if s in names]
    defs = {"species_prefixes": [""], "primary_contigs": names, "non_nuclear_contigs": [],
            "sex_chromosomes": {"_male": {s: 1 for s in sx}, "_female": {s: (2 if s.endswith("X") else 0) for s in sx}}}
    with open(os.path.join(path, "fasta", "contig-defs.json"), "w") as f:
        json.dump(defs, f, indent=4)
    with open(os.path.join(path, "genome"), "w") as f:
        f.write(genome + "\n")
