/* empty stand-in for R's <Rinternals.h>; see R.h in this directory */
