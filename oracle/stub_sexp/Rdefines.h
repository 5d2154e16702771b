/* empty stand-in for R's <Rdefines.h>; see R.h in this directory */
