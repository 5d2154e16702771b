/* empty stand-in for R's <Rmath.h>; see R.h in this directory */
