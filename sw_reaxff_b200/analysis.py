"""Host side of `fix reax/c/bonds` and `fix reax/c/species`: the file formats of the reference, fed by the device tables.

The GPU produces the connection table (Rxb.bond_table) and the molecule compositions (Rxb.species_result); what is left
for the host is what the reference's rank 0 does with the gathered buffers:
  * bonds_text      — FixReaxCBondsSunway::RecvBuffer          (fix_reaxc_bonds_sunway.cpp:264-330)
  * find_species    — FixReaxCSpeciesSunway::FindSpecies       (fix_reaxc_species_sunway.cpp:652-717)
  * species_text    — FixReaxCSpeciesSunway::WriteFormulas     (fix_reaxc_species_sunway.cpp:745-780)
The C++ twins used by the lmp_b200 driver are FixReaxCBondsB200 / FixReaxCSpeciesB200 in host/styles_b200.cpp.
"""
import numpy as np


def bonds_text(tables, ntimestep, natoms, bo_cut):
    """tables: one Rxb.bond_table() dict per rank, in rank order (the reference prints rank 0's atoms first)."""
    if isinstance(tables, dict):
        tables = [tables]
    maxnum = max(t["max_nb"] for t in tables)
    out = ["# Timestep %d \n" % ntimestep, "# \n", "# Number of particles %d \n" % natoms, "# \n",
           "# Max number of bonds per atom %d with coarse bond order cutoff %5.3f \n" % (maxnum, bo_cut),
           "# Particle connection table and bond orders \n", "# id type nb id_1...id_nb mol bo_1...bo_nb abo nlp q \n"]
    for t in tables:
        off = t["off"]
        for i in range(len(t["tag"])):
            a, b = off[i], off[i + 1]
            line = " %d %d %d" % (t["tag"][i], t["type"][i], b - a)
            line += "".join(" %d" % j for j in t["nbr"][a:b])
            line += " 0"                      # atom->molecule == NULL for atom_style charge
            line += "".join("%14.3f" % v for v in t["bo"][a:b])
            line += "%14.3f%14.3f%14.3f\n" % (t["abo"][i], t["nlp"][i], t["q"][i])
            out.append(line)
    out.append("# \n")
    return "".join(out)


def find_species(composition):
    """Unique compositions in order of first appearance and their molecule counts."""
    comp = np.asarray(composition)
    names, counts, index = [], [], {}
    for row in comp:
        key = tuple(int(v) for v in row)
        k = index.get(key)
        if k is None:
            index[key] = len(names)
            names.append(key)
            counts.append(1)
        else:
            counts[k] += 1
    return names, counts


def species_text(ntimestep, composition, elements="CHON"):
    names, counts = find_species(composition)
    out = "# Timestep     No_Moles     No_Specs     "
    for key in names:
        for j, c in enumerate(key):
            if c != 0:
                out += elements[j]
                if c != 1:
                    out += "%d" % c
        out += "\t"
    out += "\n"
    out += "%d" % ntimestep
    out += "%11d%11d\t" % (len(composition), len(names))
    for c in counts:
        out += " %d\t" % c
    out += "\n"
    return out
