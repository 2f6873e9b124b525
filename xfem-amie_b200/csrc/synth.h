// synth.h -- internal: access to the recipe behind an amie_b200_synth handle
#pragma once
#include "synth_recipe.h"
struct amie_b200_synth ;
int synth_build_recipe(SynthRecipe & R, const char * preset, int n, uint64_t seed) ;
const SynthRecipe * synth_recipe_of(const amie_b200_synth * s) ;
