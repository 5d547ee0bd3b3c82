/* include/compat/core.hpp -- stands in for inst/include/core.hpp:37-114: Hector::Core with
 * init / setData / addVisitor / prepareToRun / run / reset / shutDown / sendMessage /
 * getStartDate / getEndDate / getCurrentDate / getTrackingDate / getTrackingData / getBiomeList /
 * createBiome / deleteBiome / renameBiome / getRun_name / getGlobalLogger and the static registry
 * mkcore / getcore / delcore, over the B200 engine. */
#ifndef HECTOR_B200_COMPAT_CORE_HPP
#define HECTOR_B200_COMPAT_CORE_HPP
#ifndef HECTOR_B200_AS_HECTOR
#define HECTOR_B200_AS_HECTOR
#endif
#include "../hector_b200_core.hpp"
#endif
