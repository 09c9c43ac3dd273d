/* placeholder until the transform oracle lands */ typedef int orc_tr_placeholder;
