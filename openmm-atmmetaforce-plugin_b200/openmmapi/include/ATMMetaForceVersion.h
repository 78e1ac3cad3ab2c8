#ifndef ATMMETAFORCE_VERSION_H_
#define ATMMETAFORCE_VERSION_H_
/* same version id the reference reports (ref: openmmapi/include/ATMMetaForceVersion.h:4) */
#define ATMMETAFORCE_VERSION "0.3.1"
#endif
