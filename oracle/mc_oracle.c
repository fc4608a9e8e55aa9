int oracle_mc_placeholder(void){return 0;}
