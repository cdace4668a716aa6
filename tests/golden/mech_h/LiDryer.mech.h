#define n_species 9
#define n_active_species 8
#define species_names_length 29
#define n_reactions 21
__KINETIX_CONST__ char species_names[species_names_length] = "H2 O2 H2O H O OH HO2 H2O2 N2";
__KINETIX_CONST__ cfloat kinetix_molar_mass[n_species] = {0.002016, 0.031998, 0.018015, 0.001008, 0.015999, 0.017006999999999998, 0.033006, 0.034013999999999996, 0.028014};
__KINETIX_CONST__ cfloat kinetix_rcp_molar_mass[n_species] = {496.031746031746, 31.251953247077942, 55.50929780738274, 992.063492063492, 62.503906494155885, 58.79931792791204, 30.297521662727988, 29.39965896395602, 35.69643749553794};
#define __KINETIX_NSPECIES__ n_species
#define __KINETIX_NACTIVESPECIES__ n_active_species